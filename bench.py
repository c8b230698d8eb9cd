#!/usr/bin/env python
"""bench.py -- BASELINE config 2: 30-qubit complex128 RX/RY/RZ + CNOT-ring gate layers.

One "step" = LAYERS (default 4) layers of [RX,RY,RZ on every wire + CNOT ring] = 120 gates per layer
applied to a device-resident 2^n state (n = 30 on one GPU: 17.2 GB >> the 126 MB L2, so no L2
flush is needed between steps).  The state is prepared by one Hadamard layer (untimed).

Metric (both arms): "gate-layer GB/s" = reference-equivalent bytes per second, i.e. the bytes the
reference's one-gate-per-sweep schedule (reference StateVectorKokkos.hpp:807-824) has to move for
the same gates, sum_g touch(g) * 2*16*2^n with touch = 1 for RX/RY/RZ and 1/2 for CNOT
(SURVEY.md section 8d), divided by the measured time.  Because the engine fuses many gates per HBM
pass this exceeds the HBM peak; the `roofline` object reports what the tile kernel really moves:
(sweeps executed) * 2*16*2^n bytes / time, against MEASURED_PEAKS.json.

  value     b2sv_apply_ops on a pre-built op-list handle with the state resident in HBM (includes
            the host-side fusion scheduling and the few-KB pass-descriptor upload), CUDA events on
            the engine's stream.
  e2e       the call a user of the reference API makes: LightningKokkos_C128.apply(names, wires,
            inverses, params) from HOST Python lists (marshalling, lowering, scheduling, descriptor
            H2D) followed by ExpectationValue("PauliZ") read back to the host (D2H of the scalar).
            The state vector itself is device-resident by the reference's API contract
            (HostToDevice/DeviceToHost are explicit calls, StateVectorKokkos.hpp:1618-1628).
  adjoint_jacobian
            (N = 1 only) the second half of BASELINE.json's metric: BASELINE config 3, a 24-qubit
            hardware-efficient ansatz with 504 parameters and a 100-term Pauli Hamiltonian, seconds
            per adjoint Jacobian through AdjointJacobianKokkos_C128.adjoint_jacobian.
  cpu_baseline / --impl reference
            the UNMODIFIED reference functors (oracle/_ref/libref_oracle.so: reference headers over
            the OpenMP Kokkos stand-in) on the host cores, on a bounded sample of the same layer.

Multi-GPU (torchrun, one rank per GPU): weak scaling, n = 30 + log2(N) qubits, rank = top index
bits; gates on global qubits trigger NVLink qubit swaps (csrc/comm.cpp).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOUCH = {"RX": 1.0, "RY": 1.0, "RZ": 1.0, "CNOT": 0.5}
NCU_TRAFFIC_30Q = 34.328e9  # dram read + write bytes per tile-kernel launch (profiles/r1_ncu_tile_v20.summary.txt)


def layer_circuit(n, layers, seed=42):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(layers):
        for w in range(n):
            for g in ("RX", "RY", "RZ"):
                ops.append((g, [w], False, [float(rng.uniform(0, 2 * np.pi))]))
        for w in range(n):
            ops.append(("CNOT", [w, (w + 1) % n], False, []))
    return ops


def ref_equiv_bytes(ops, n, amp_bytes=16):
    sweep = 2.0 * amp_bytes * (1 << n)
    return sum(TOUCH[o[0]] for o in ops) * sweep


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation (its functors over the OpenMP Kokkos stand-in)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference",
                          "unavailable": "oracle/_ref/libref_oracle.so missing (needs the build container)"}))
        return 0
    n = args.qubits
    cores = ref.num_threads()
    # bounded sample of one layer: RX,RY,RZ on a low / middle / high wire + 3 CNOTs of the ring
    rng = np.random.default_rng(42)
    sample = []
    for w in (0, n // 2, n - 1):
        for g in ("RX", "RY", "RZ"):
            sample.append((g, [w], False, [float(rng.uniform(0, 2 * np.pi))]))
    for w in (0, n // 2, n - 1):
        sample.append(("CNOT", [w, (w + 1) % n], False, []))
    sv = ref.RefStateVector(n, np.complex128)
    for w in (0, n - 1):
        sv.apply("Hadamard", [w])
    nbytes = ref_equiv_bytes(sample, n)
    for _ in range(args.warmup):
        sv.apply_ops(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sv.apply_ops(sample)
    dt = (time.perf_counter() - t0) / args.steps
    val = nbytes / dt / 1e9
    desc = (f"{len(sample)} gates of one layer (RX,RY,RZ on wires 0,{n // 2},{n - 1} + 3 ring CNOTs) "
            f"on the full {n}-qubit c128 state, {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "30q c128 gate-layer GB/s", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": {"workload": f"{n}-qubit c128 RX/RY/RZ + CNOT-ring layers (BASELINE config 2)",
                   "qubits": n, "sample": desc,
                   "bytes": "reference-equivalent: sum_g touch(g)*2*16*2^n"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": desc},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_sample(n, budget_s=20.0):
    """Rank 0, N=1 only: time the reference on a bounded sample (see run_reference)."""
    try:
        from oracle import ref
        if not ref.available():
            return {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                    "sample": "oracle/_ref missing"}
        cores = ref.num_threads()
        rng = np.random.default_rng(42)
        sample = []
        for w in (0, n // 2, n - 1):
            for g in ("RX", "RY", "RZ"):
                sample.append((g, [w], False, [float(rng.uniform(0, 2 * np.pi))]))
        for w in (0, n // 2, n - 1):
            sample.append(("CNOT", [w, (w + 1) % n], False, []))
        sv = ref.RefStateVector(n, np.complex128)
        sv.apply("Hadamard", [0])
        sv.apply("Hadamard", [n - 1])
        t0 = time.perf_counter()
        reps = 0
        while True:
            sv.apply_ops(sample)
            reps += 1
            if time.perf_counter() - t0 > budget_s or reps >= 3:
                break
        dt = (time.perf_counter() - t0) / reps
        return {"value": ref_equiv_bytes(sample, n) / dt / 1e9, "unit": "GB/s", "cores": cores,
                "kind": "reference",
                "sample": f"{len(sample)} gates of one layer (RX,RY,RZ on wires 0,{n // 2},{n - 1} + 3 "
                          f"ring CNOTs) on the full {n}-qubit c128 state, {reps} repetition(s), "
                          f"{dt:.2f} s each, reference functors over the OpenMP Kokkos stand-in"}
    except Exception as e:  # the baseline must never break the bench line
        return {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                "sample": f"failed: {e}"}


def adjoint_sample(ops_module):
    """The second half of BASELINE.json's metric ("adjoint-Jacobian s/circuit"): BASELINE config 3
    (24 qubits, 504 parameters, 100-term Pauli Hamiltonian) through the binding's adjoint_jacobian,
    single GPU. Never allowed to break the bench line."""
    try:
        bdir = os.path.join(ROOT, "benchmarks")
        if bdir not in sys.path:
            sys.path.insert(0, bdir)
        import configs as cfgs
        r = cfgs.config3(ops_module, 3)
        return {"s_per_jacobian": r["s_per_jacobian"], "s_forward": r["s_forward"],
                "workload": r["workload"], "jac_norm": r["jac_norm"]}
    except Exception as e:  # pragma: no cover
        return {"s_per_jacobian": None, "error": str(e)[:200]}


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

    g = world.bit_length() - 1
    assert (1 << g) == world, "number of GPUs must be a power of two"
    n = args.qubits + g  # weak scaling: 2^qubits amplitudes per GPU
    circ = layer_circuit(n, args.layers, seed=42)
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]

    if world == 1:
        sv = ops.LightningKokkos_C128(n)
    else:
        from pennylane_lightning_kokkos_b200 import dist as b2dist
        sv = b2dist.create_sharded_state(ops, n, np.complex128, local_rank)
    had = ops.OpsStructKokkos_C128(["Hadamard"] * n, [[] for _ in range(n)],
                                   [[w] for w in range(n)], [False] * n)
    sv.apply_ops(had)
    oplist = ops.OpsStructKokkos_C128(names, params, wires, invs)
    stream = torch.cuda.ExternalStream(sv.stream_ptr(), device=torch.device("cuda", local_rank))

    def barrier():
        sv.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        sv.sync()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, t0, t1

    # ---- device-resident arm: pre-built op list handle
    def step_dev():
        sv.apply_ops(oplist)

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sv.reset_stats()
    ms_dev, t0, t1 = timed(step_dev, args.steps)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    st = sv.stats()
    sweeps_per_step = st["sweeps"] / args.steps
    launches = st["launches"]
    comm = sv.comm_stats() if hasattr(sv, "comm_stats") else {}

    # ---- end-to-end arm: host lists -> apply -> expval read back
    def step_e2e():
        sv.apply(names, wires, invs, params)
        return sv.ExpectationValue("PauliZ", [n - 1], [], np.zeros(0))

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    tw0 = time.perf_counter()
    ez = 0.0
    for _ in range(args.steps):
        ez = step_e2e()
    sv.sync()
    tw1 = time.perf_counter()
    ms_e2e = (tw1 - tw0) * 1e3
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    blob_bytes = sv.last_upload_bytes() if hasattr(sv, "last_upload_bytes") else 0
    norm = sv.ExpectationValue("Identity", [0], [], np.zeros(0))

    nbytes = ref_equiv_bytes(circ, n)  # whole job (all ranks)
    value = nbytes / (ms_dev / args.steps * 1e-3) / 1e9
    e2e_value = nbytes / (ms_e2e / args.steps * 1e-3) / 1e9
    peaks, which = measured_peaks()
    sweep_bytes = 2.0 * 16 * (1 << args.qubits)  # per GPU, per launch of the tile kernel
    tile_launches = st["sweeps"]
    avg_launch_ms = ms_dev / max(1, tile_launches) if not comm.get("swap_ms") else None
    achieved = sweep_bytes / (ms_dev / max(1, tile_launches) * 1e-3) / 1e9
    line = None
    if rank == 0:
        cpu = cpu_baseline_sample(args.qubits) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": "30q c128 gate-layer GB/s", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
            "config": {
                "workload": f"{n}-qubit c128 RX/RY/RZ + CNOT-ring layers (BASELINE config 2), "
                            f"{args.layers} layers = {len(circ)} gates per step",
                "qubits": n, "qubits_per_gpu": args.qubits, "layers_per_step": args.layers,
                "gates_per_step": len(circ),
                "bytes": "reference-equivalent: sum_g touch(g)*2*16*2^n (touch 1 for RX/RY/RZ, 1/2 for CNOT)",
                "l2": f"state {16 * (1 << args.qubits) / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                "sweeps_per_step": sweeps_per_step,
                "gates_per_s": len(circ) / (ms_dev / args.steps * 1e-3),
                "parallelism": f"state sharded over {world} GPU(s), rank = top {g} index bits",
                "checks": {"norm": norm, "expval_Z_last": ez},
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                # (profiles/r1_ncu_tile_v20.summary.txt at n = 30)
                "traffic": NCU_TRAFFIC_30Q if args.qubits == 30 else None, "peak_source": which,
                "kernel": "tile_exec_kernel<double,12,4,256,2,3,FACT> (persistent, 148 CTAs x 640 threads)",
                "algorithmic_bytes_per_launch": sweep_bytes,
                "avg_launch_ms": avg_launch_ms,
                "note": "achieved = 2*16*2^n bytes per tile-kernel launch / (CUDA-event time of the "
                        "timed region / tile-kernel launches); the region holds only tile-kernel "
                        "launches (pass descriptors travel as kernel parameters). Sustained figure: "
                        "the region is steps x ~0.1 s of back-to-back launches under the 1 kW power cap "
                        "(a single launch under ncu runs 13 % faster, profiles/r1_ncu_tile_v20.summary.txt).",
            },
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": blob_bytes,
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
                    "call": "LightningKokkos_C128.apply(names, wires, inverses, params) from host "
                            "lists + ExpectationValue('PauliZ') read back; state device-resident"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if comm:
            line["comm"] = comm
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_adjoint:
            line["adjoint_jacobian"] = adjoint_sample(ops)
        print(json.dumps(line))
    del sv
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=30, help="qubits per GPU (weak scaling)")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adjoint", action="store_true", help="skip the config-3 adjoint Jacobian sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
