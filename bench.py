#!/usr/bin/env python
"""bench.py -- BASELINE config 2: 30-qubit complex128 RX/RY/RZ + CNOT-ring gate layers.

One "step" = LAYERS (default 4) layers of [RX,RY,RZ on every wire + CNOT ring] = 120 gates per layer
applied to a device-resident 2^n state (n = 30 on one GPU: 17.2 GB >> the 126 MB L2, so no L2
flush is needed between steps).  The state is prepared by one Hadamard layer (untimed).

Metric (both arms): gate-layer throughput in reference-equivalent GB/s, i.e. the bytes the
reference's one-gate-per-sweep schedule (reference StateVectorKokkos.hpp:807-824) moves for the
same gates, sum_g touch(g) * 2*16*2^n with touch = 1 for RX/RY/RZ and 1/2 for CNOT (SURVEY.md
section 8d), divided by the measured time.  For the reference arm this IS its memory traffic; the
engine fuses many gates per HBM pass, so its figure exceeds the HBM peak -- what the tile kernel
really moves is in `roofline` (algorithmic bytes of one pass / the pass's measured duration,
against MEASURED_PEAKS.json) and repeated at top level as `hbm_gbs` / `hbm_frac`.

  value     b2sv_apply_ops on a pre-built op-list handle with the state resident in HBM (includes
            the host-side fusion scheduling and the pass-descriptor upload), CUDA events on the
            engine's stream, max over ranks.
  roofline  every tile-kernel launch of two further steps is bracketed by its own pair of CUDA events
            (b2sv_trace_begin/_end): `achieved` = 2*16*2^n bytes / mean launch duration; at N > 1 the
            exchanges are timed the same way and reported against the NVLink peer-copy figure.
  e2e       the call a user of the reference API makes: LightningKokkos_C128.apply(names, wires,
            inverses, params) from HOST Python lists (marshalling, lowering, scheduling, descriptor
            H2D) followed by ExpectationValue("PauliZ") read back to the host (D2H of the scalar).
            The state vector itself is device-resident by the reference's API contract
            (HostToDevice/DeviceToHost are explicit calls, StateVectorKokkos.hpp:1618-1628).
  parity    checked in the same run.  N = 1: the reference (oracle/_ref) applies one full layer to a
            30-qubit |0..0> on the host cores (that is also the cpu_baseline timing), the engine does
            the same on the GPU and 4096 sampled amplitudes + the norm are compared.  N > 1: 17..19
            qubit layered and random circuits on sharded states against the NumPy oracle.
  adjoint_jacobian
            (N = 1) the second half of BASELINE.json's metric: BASELINE config 3, a 24-qubit
            hardware-efficient ansatz with 504 parameters and a 100-term Pauli Hamiltonian, seconds
            per adjoint Jacobian; the same ansatz at 20 qubits is run through the reference on the
            host (cpu_baseline + parity of the full Jacobian).
  --impl reference
            the UNMODIFIED reference functors (oracle/_ref/libref_oracle.so: reference headers over
            the OpenMP Kokkos stand-in) on all host cores; one step = one full 120-gate layer of the
            same circuit on a 30-qubit state.

Multi-GPU (torchrun, one rank per GPU): weak scaling, n = qubits + log2(N), rank = top index
bits; gates on global qubits trigger NVLink exchanges (csrc/comm.cpp).
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

TOUCH = {"RX": 1.0, "RY": 1.0, "RZ": 1.0, "CNOT": 0.5, "CRX": 0.5, "CRY": 0.5, "CRZ": 0.5, "CRot": 0.5,
         "IsingXX": 1.0, "IsingYY": 1.0, "IsingZZ": 1.0, "MultiRZ": 1.0, "SingleExcitation": 0.5,
         "DoubleExcitation": 0.125, "Toffoli": 0.25, "ControlledPhaseShift": 0.25}
METRIC = ("30q c128 gate-layer GB/s (reference-equivalent bytes: what one gate per HBM sweep moves "
          "for the same gates; roofline.achieved is the HBM traffic rate vs peak)")
NVLINK_PEER_GBS = 770.0  # measured peer-copy GB/s per direction, B200_PROFILING.md
PARITY_SAMPLES = 4096


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


def controlled_circuit(n, layers, seed=42):
    """--workload controlled: the controlled / parametric two- to four-qubit families of north_star on
    neighbouring wires (CRX, CRot, IsingXX, SingleExcitation, DoubleExcitation, MultiRZ, Toffoli)."""
    rng = np.random.default_rng(seed)
    u = lambda: float(rng.uniform(0, 2 * np.pi))
    ops = []
    for _ in range(layers):
        for w in range(0, n - 1, 2):
            ops.append(("CRX", [w, w + 1], False, [u()]))
        for w in range(1, n - 1, 2):
            ops.append(("CRot", [w + 1, w], False, [u(), u(), u()]))
        for w in range(0, n - 1, 2):
            ops.append(("IsingXX", [w, w + 1], False, [u()]))
        for w in range(1, n - 1, 2):
            ops.append(("SingleExcitation", [w, w + 1], False, [u()]))
        for w in range(0, n - 3, 4):
            ops.append(("DoubleExcitation", [w, w + 1, w + 2, w + 3], False, [u()]))
        for w in range(0, n - 2, 3):
            ops.append(("MultiRZ", [w, w + 1, w + 2], False, [u()]))
        for w in range(0, n - 2, 3):
            ops.append(("Toffoli", [w + 2, w, w + 1], False, []))
    return ops


def ref_equiv_bytes(ops, n, amp_bytes=16):
    sweep = 2.0 * amp_bytes * (1 << n)
    return sum(TOUCH[o[0]] for o in ops) * sweep


def static_config(world, qubits_per_gpu, layers):
    """The workload description: identical in both arms (the reference arm samples it)."""
    g = world.bit_length() - 1
    n = qubits_per_gpu + g
    return {
        "workload": f"{n}-qubit c128 RX/RY/RZ + CNOT-ring layers (BASELINE config 2), "
                    f"{layers} layers = {4 * n * layers} gates per step",
        "qubits": n, "qubits_per_gpu": qubits_per_gpu, "layers_per_step": layers,
        "gates_per_step": 4 * n * layers,
        "bytes": "reference-equivalent: sum_g touch(g)*2*16*2^n (touch 1 for RX/RY/RZ, 1/2 for CNOT)",
        "l2": f"state {16 * (1 << qubits_per_gpu) / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
        "parallelism": f"state sharded over {world} GPU(s), rank = top {g} index bits",
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(qubits_per_gpu):
    """dram read + write bytes of one tile-kernel launch from the committed ncu capture of this
    round (profiles/r2_ncu_tile.json, written by profiles/summarize_ncu.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_tile.json")
    try:
        with open(p) as f:
            d = json.load(f)
        if int(d.get("qubits", -1)) == qubits_per_gpu:
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        pass
    return None


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
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_reference():
    """oracle/_ref with every host core, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1); must run before libgomp is initialised by the oracle's first call."""
    nthr = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthr)
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ.setdefault("OMP_PLACES", "cores")
    from oracle import ref
    if not ref.available():
        return None, 0
    ref.set_num_threads(nthr)
    return ref, ref.num_threads()


def sample_indices(n, count=PARITY_SAMPLES, seed=7):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 1 << n, size=count, dtype=np.uint64)
    idx[:4] = (0, 1, (1 << n) - 1, 1 << (n - 1))
    return idx


def reference_layer_run(ref, n, steps, warmup, keep_amplitudes):
    """One step = the first full layer (120 gates at n = 30) of the bench circuit on an n-qubit c128
    state held by the reference.  Returns (seconds per step, sampled amplitudes after the FIRST
    application to |0..0> or None, description)."""
    layer = layer_circuit(n, 1, seed=42)
    sv = ref.RefStateVector(n, np.complex128)
    amps = None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        sv.apply_ops(layer)
        dt = time.perf_counter() - t0
        if it == 0 and keep_amplitudes:
            amps = sv.amplitudes(sample_indices(n).astype(np.int64))
        if it >= warmup:
            times.append(dt)
    desc = (f"one full layer of the bench circuit ({len(layer)} gates: RX,RY,RZ on every wire + CNOT "
            f"ring) per step on a {n}-qubit c128 state, reference functors over the OpenMP Kokkos "
            f"stand-in, {len(times)} timed step(s)")
    return float(np.mean(times)), amps, desc, layer


def run_reference(args):
    """The reference's own CPU implementation (its functors over the OpenMP Kokkos stand-in)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref, cores = load_reference()
    if ref is None:
        print(json.dumps({"impl": "reference",
                          "unavailable": "oracle/_ref/libref_oracle.so missing (needs the build container)"}))
        return 0
    n = args.ref_qubits
    dt, _, desc, layer = reference_layer_run(ref, n, args.steps, args.warmup, False)
    val = ref_equiv_bytes(layer, n) / dt / 1e9
    sample = desc + f"; {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": static_config(world, args.qubits, args.layers),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_and_parity(sv, n):
    """Rank 0, N = 1: the reference applies one full layer to |0..0> on the host (timed = the
    cpu_baseline), the engine applies the same layer on the GPU; sampled amplitudes are compared."""
    try:
        ref, cores = load_reference()
        if ref is None:
            return ({"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                     "sample": "oracle/_ref missing"}, None)
        dt, amps_ref, desc, layer = reference_layer_run(ref, n, 1, 0, True)
        cpu = {"value": ref_equiv_bytes(layer, n) / dt / 1e9, "unit": "GB/s", "cores": cores,
               "kind": "reference", "sample": desc + f", {dt:.1f} s; {cores} OpenMP threads"}
        sv.setBasisState(0)
        sv.apply([c[0] for c in layer], [c[1] for c in layer], [c[2] for c in layer],
                 [c[3] for c in layer])
        amps = sv.amplitudes(sample_indices(n))
        norm = sv.ExpectationValue("Identity", [0], [], np.zeros(0))
        scale = float(np.max(np.abs(amps_ref)))
        err = float(np.max(np.abs(amps - amps_ref)) / scale)
        parity = {"max_rel_err": max(err, abs(norm - 1.0)), "amplitude_rel_err": err,
                  "norm_err": abs(norm - 1.0), "n": n, "world": 1, "samples": int(amps.size),
                  "tolerance": 1e-12, "ok": bool(err < 1e-12 and abs(norm - 1.0) < 1e-12),
                  "against": "oracle/_ref (reference functors) on the same 120-gate layer from |0..0>"}
        return cpu, parity
    except Exception as e:  # the baseline must never break the bench line
        return ({"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                 "sample": f"failed: {e}"}, {"ok": False, "error": str(e)[:200]})


def sharded_parity(ops, b2dist, dist, torch, local_rank, world):
    """N > 1, before timing: layered + random circuits on sharded states vs the NumPy oracle."""
    from cases import layered_circuit, random_circuit
    from oracle import np_oracle as npo

    g = world.bit_length() - 1
    worst, cases = 0.0, []
    for n, circ, tag in ((16 + g, layered_circuit(16 + g, 2, seed=3), "layers"),
                         (13 + 2 * g, random_circuit(13 + 2 * g, 120, seed=5), "random"),
                         (19 + g, layered_circuit(19 + g, 3, seed=8), "layers, sliced + overlapped")):
        # the last case forces the pipelined path (shard cut into slices, exchanges on a second
        # stream beside the passes) that the timed 30-qubit-per-GPU run takes
        if "sliced" in tag:
            os.environ["B2SV_PIPE_MIN_SUB"] = "0"
        else:
            os.environ.pop("B2SV_PIPE_MIN_SUB", None)
        sv = b2dist.create_sharded_state(ops, n, np.complex128, local_rank)
        sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ], [c[3] for c in circ])
        ez = [sv.ExpectationValue("PauliZ", [w], [], np.zeros(0)) for w in (0, n - 1)]
        mine = np.zeros(1 << (n - g), dtype=np.complex128)
        sv.DeviceToHost(mine)
        t = torch.from_numpy(mine.view(np.float64)).cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        full = np.concatenate([p.cpu().numpy() for p in parts]).view(np.complex128)
        psi0 = np.zeros(1 << n, dtype=complex)
        psi0[0] = 1
        want = npo.apply_ops(psi0, n, circ)
        err = float(np.max(np.abs(full - want)) / np.max(np.abs(want)))
        ez_want = [npo.expval(want, n, ("named", "PauliZ", [w])) for w in (0, n - 1)]
        err = max(err, max(abs(a - b) for a, b in zip(ez, ez_want)))
        worst = max(worst, err)
        cases.append(f"{tag} n={n}: {err:.1e}")
        del sv
        dist.barrier()
    os.environ.pop("B2SV_PIPE_MIN_SUB", None)
    return {"max_rel_err": worst, "n": 16 + g, "world": world, "tolerance": 1e-12,
            "ok": bool(worst < 1e-12), "cases": cases,
            "against": "oracle/np_oracle.py (full state gathered from all ranks + <Z> on a global and a local wire)"}


def sharded_adjoint_sample(ops, b2dist, local_rank, world, dist, q_per_gpu=24, layers=7, terms=100):
    """BASELINE config 3's ansatz on a state sharded over `world` GPUs (24 qubits per GPU): seconds
    per adjoint Jacobian, max over ranks, plus a parity check of the same code path at 12 + g qubits
    against the NumPy oracle."""
    import time as _t

    from cases import random_pauli_hamiltonian
    from oracle import np_oracle as npo

    bdir = os.path.join(ROOT, "benchmarks")
    if bdir not in sys.path:
        sys.path.insert(0, bdir)
    import configs as cfgs
    g = world.bit_length() - 1

    def build(n, layers, terms):
        circ = cfgs.hea_circuit(n, layers)
        ham = random_pauli_hamiltonian(n, terms, seed=42)
        tobs = []
        for _, word in ham:
            fac = [ops.NamedObsKokkos_C128(l, [w]) for l, w in word]
            tobs.append(fac[0] if len(fac) == 1 else ops.TensorProdObsKokkos_C128(fac))
        H = ops.HamiltonianKokkos_C128(np.array([c for c, _ in ham]), tobs)
        names, wires, invs, params = cfgs.split(circ)
        sv = b2dist.create_sharded_state(ops, n, np.complex128, local_rank)
        sv.apply(names, wires, invs, params)
        adj = ops.AdjointJacobianKokkos_C128()
        oplist = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs,
                                     [np.zeros(0, dtype=complex) for _ in names])
        tp = list(range(sum(1 for p in params if len(p))))
        return circ, ham, sv, adj, oplist, H, tp

    # parity at a size the oracle finishes in seconds
    n_s = 13 + 2 * g  # a shard must hold at least tile_bits + log2(world) qubits
    circ, ham, sv, adj, oplist, H, tp = build(n_s, 2, 12)
    jac = adj.adjoint_jacobian(sv, [H], oplist, tp)
    psi0 = np.zeros(1 << n_s, dtype=complex)
    psi0[0] = 1
    final = npo.apply_ops(psi0, n_s, circ)
    hob = ("hamiltonian", [c for c, _ in ham],
           [("tensor", [("named", l, [w]) for l, w in word]) for _, word in ham])
    want = npo.adjoint_jacobian(final, n_s, [hob], circ, tp)
    err = float(np.max(np.abs(jac - want)) / max(np.max(np.abs(want)), 1e-300))
    del sv
    dist.barrier()
    n = q_per_gpu + g
    circ, ham, sv, adj, oplist, H, tp = build(n, layers, terms)
    adj.adjoint_jacobian(sv, [H], oplist, tp)
    sv.sync()
    dist.barrier()
    t0 = _t.perf_counter()
    jac = adj.adjoint_jacobian(sv, [H], oplist, tp)
    dt = _t.perf_counter() - t0
    import torch
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del sv
    dist.barrier()
    return {"s_per_jacobian": float(t.item()), "jac_norm": float(np.linalg.norm(jac)),
            "workload": f"{n}q HEA {layers} layers, {len(tp)} params, adjoint Jacobian of a {terms}-term Pauli "
                        f"Hamiltonian, c128, state sharded over {world} GPUs ({q_per_gpu} qubits per GPU)",
            "parity": {"max_rel_err": err, "n": n_s, "world": world, "tolerance": 1e-12, "ok": bool(err < 1e-12),
                       "against": "oracle/np_oracle.py adjoint_jacobian (2 layers, 12-term Hamiltonian)"}}


def adjoint_sample(ops_module):
    """The second half of BASELINE.json's metric ("adjoint-Jacobian s/circuit"): BASELINE config 3
    (24 qubits, 504 parameters, 100-term Pauli Hamiltonian) through the binding's adjoint_jacobian on
    one GPU, plus the same ansatz at 20 qubits through the reference on the host cores (cpu baseline
    and parity of the full Jacobian). Never allowed to break the bench line."""
    try:
        bdir = os.path.join(ROOT, "benchmarks")
        if bdir not in sys.path:
            sys.path.insert(0, bdir)
        import configs as cfgs
        r = cfgs.config3(ops_module, 3)
        out = {"s_per_jacobian": r["s_per_jacobian"], "s_forward": r["s_forward"],
               "workload": r["workload"], "jac_norm": r["jac_norm"], "roofline": r.get("roofline")}
        try:
            out["parity_20q"] = cfgs.config3_vs_reference(ops_module, n=20, layers=7, terms=100)
        except Exception as e:  # pragma: no cover
            out["parity_20q"] = {"ok": False, "error": str(e)[:200]}
        return out
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
    c64 = args.dtype == "c64"
    np_dtype = np.complex64 if c64 else np.complex128
    amp_bytes = 8 if c64 else 16
    SV = ops.LightningKokkos_C64 if c64 else ops.LightningKokkos_C128
    OPS = ops.OpsStructKokkos_C64 if c64 else ops.OpsStructKokkos_C128
    circ = (controlled_circuit if args.workload == "controlled" else layer_circuit)(n, args.layers, seed=42)
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]

    parity = None
    if world > 1:
        from pennylane_lightning_kokkos_b200 import dist as b2dist
        if not args.no_parity:
            try:
                parity = sharded_parity(ops, b2dist, dist, torch, local_rank, world)
            except Exception as e:
                parity = {"ok": False, "error": str(e)[:300], "world": world}
        sv = b2dist.create_sharded_state(ops, n, np_dtype, local_rank)
    else:
        sv = SV(n)
    had = OPS(["Hadamard"] * n, [[] for _ in range(n)], [[w] for w in range(n)], [False] * n)
    sv.apply_ops(had)
    oplist = OPS(names, params, wires, invs)
    stream = torch.cuda.ExternalStream(sv.stream_ptr(), device=torch.device("cuda", local_rank))

    def barrier():
        sv.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

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
        ms = reduce_max(e0.elapsed_time(e1))
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
    comm0 = sv.comm_stats() if world > 1 else {}
    ms_dev, t0, t1 = timed(step_dev, args.steps)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    st = sv.stats()
    sweeps_per_step = st["sweeps"] / args.steps
    launches = st["launches"]
    comm = sv.comm_stats() if world > 1 else {}

    # ---- per-kernel timing: every launch of two more steps between its own pair of CUDA events
    barrier()
    sv.trace_begin()
    trace_steps = 2
    for _ in range(trace_steps):
        step_dev()
    trace = sv.trace_end()
    pass_ms = [d for k, _, d in trace if k == 0]
    xchg_ms = [d for k, _, d in trace if k == 2]
    step_ms_traced = (max(s + d for _, s, d in trace) - min(s for _, s, _ in trace)) / trace_steps if trace else None
    barrier()

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
    ms_e2e = reduce_max((tw1 - tw0) * 1e3)
    blob_bytes = sv.last_upload_bytes() if hasattr(sv, "last_upload_bytes") else 0
    norm = sv.ExpectationValue("Identity", [0], [], np.zeros(0))

    nbytes = ref_equiv_bytes(circ, n, amp_bytes)  # whole job (all ranks)
    value = nbytes / (ms_dev / args.steps * 1e-3) / 1e9
    e2e_value = nbytes / (ms_e2e / args.steps * 1e-3) / 1e9
    peaks, which = measured_peaks()
    sweep_bytes = 2.0 * amp_bytes * (1 << args.qubits)  # per GPU, per launch of the tile kernel
    avg_pass_ms = float(np.mean(pass_ms)) if pass_ms else None
    achieved = sweep_bytes / (avg_pass_ms * 1e-3) / 1e9 if avg_pass_ms else None
    # ---- N > 1: the adjoint half of the metric on a sharded state (24 qubits per GPU, collective)
    adj_sharded = None
    if world > 1 and not args.no_adjoint and args.workload == "layers" and not c64:
        try:
            adj_sharded = sharded_adjoint_sample(ops, b2dist, local_rank, world, dist)
        except Exception as e:
            adj_sharded = {"s_per_jacobian": None, "error": str(e)[:300]}
    line = None
    if rank == 0:
        cpu = None
        default_workload = args.workload == "layers" and not c64
        if world == 1 and not args.no_cpu_baseline and default_workload:
            cpu, parity = cpu_baseline_and_parity(sv, n)
        cfg = static_config(world, args.qubits, args.layers)
        if not default_workload:
            cfg["workload"] = (f"{n}-qubit {args.dtype} --workload {args.workload}, {args.layers} layers = "
                               f"{len(circ)} gates per step (not the BASELINE headline config)")
            cfg["gates_per_step"] = len(circ)
        roof = {
            "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"] if achieved else None,
            "traffic": ncu_traffic(args.qubits), "peak_source": which,
            "kernel": "tile_exec_kernel<double,12,4,...> (persistent, one CTA per SM)",
            "algorithmic_bytes_per_launch": sweep_bytes,
            "avg_launch_ms": avg_pass_ms,
            "launches_timed": len(pass_ms),
            "kernel_share_of_step": (sum(pass_ms) / trace_steps) / step_ms_traced if step_ms_traced else None,
            "note": "achieved = 2*16*2^n bytes / mean duration of the tile-kernel launches of "
                    f"{trace_steps} steps, each launch bracketed by its own CUDA events on the engine's "
                    "stream (b2sv_trace_begin/_end), back to back under the power cap; sustained-step "
                    "figure = sweeps_per_step * bytes / ms_per_step in `hbm_gbs_step`.",
        }
        hbm_step = sweeps_per_step * sweep_bytes / (ms_dev / args.steps * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c64 (f32)" if c64 else "c128 (f64)", "data": "synthetic",
            "config": cfg,
            "hbm_gbs": achieved, "hbm_frac": roof["frac"], "hbm_gbs_step": hbm_step,
            "schedule": {"sweeps_per_step": sweeps_per_step,
                         "gates_per_s": len(circ) / (ms_dev / args.steps * 1e-3),
                         "checks": {"norm": norm, "expval_Z_last": ez}},
            "roofline": roof,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": blob_bytes,
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
                    "call": "LightningKokkos_C128.apply(names, wires, inverses, params) from host "
                            "lists + ExpectationValue('PauliZ') read back; state device-resident"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world > 1:
            nx = (comm.get("swaps", 0) - comm0.get("swaps", 0)) / args.steps
            xb = (comm.get("swap_bytes_per_rank", 0) - comm0.get("swap_bytes_per_rank", 0)) / args.steps
            x_ms = sum(xchg_ms) / trace_steps if xchg_ms else 0.0
            x_gbs = xb / (x_ms * 1e-3) / 1e9 if x_ms > 0 else None
            line["comm"] = dict(comm, exchanges_per_step=nx, bytes_per_rank_per_step=xb,
                                exchange_ms_per_step=x_ms, pass_ms_per_step=sum(pass_ms) / trace_steps,
                                traced_step_ms=step_ms_traced)
            line["roofline_nvlink"] = {
                "bound": "nvlink", "achieved": x_gbs, "peak": NVLINK_PEER_GBS, "unit": "GB/s per direction",
                "frac": x_gbs / NVLINK_PEER_GBS if x_gbs else None,
                "note": "bytes each rank sends per step / time of the exchange launches of the traced steps"}
            # the step against max(HBM term, NVLink term) if the two overlapped perfectly
            hbm_term = sweeps_per_step * sweep_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
            nvl_term = xb / (NVLINK_PEER_GBS * 1e9) * 1e3
            line["roofline_step"] = {"hbm_ms": hbm_term, "nvlink_ms": nvl_term,
                                     "frac_of_max": max(hbm_term, nvl_term) / (ms_dev / args.steps),
                                     "frac_of_sum": (hbm_term + nvl_term) / (ms_dev / args.steps)}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_adjoint and default_workload:
            line["adjoint_jacobian"] = adjoint_sample(ops)
        if adj_sharded is not None:
            line["adjoint_jacobian"] = adj_sharded
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
    ap.add_argument("--ref-qubits", type=int, default=30, help="state size of the reference arm's sample")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--workload", default="layers", choices=["layers", "controlled"],
                    help="layers = BASELINE config 2 (default); controlled = CRX/CRot/Ising/excitation/MultiRZ layers")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-adjoint", action="store_true", help="skip the config-3 adjoint Jacobian sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
