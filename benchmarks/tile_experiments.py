#!/usr/bin/env python
"""Developer experiment driver: times BASELINE config 2 (30 q c128 layers) under different tile
executor / scheduler knobs, one subprocess per setting (the knobs are read once per process).
Prints one JSON line per setting: ms per step, sweeps, ms per sweep, achieved GB/s per sweep and,
with B2SV_TILE_PROF=1, the phase-timer breakdown (fractions of the lead threads' lifetime).

Knobs (all read once per process): B2SV_MAX_HEAVY (arithmetic ops per pass, default automatic),
B2SV_FACTOR (0: no factored rounds), B2SV_FUSE_STORE (0: store phase only), B2SV_TILE_LOW (contiguous
low bits of a tile, 2..6, default 5), B2SV_TILE_FLAGS (1: lockstep worker groups, 2: free-running),
B2SV_TILE_PROF (1: phase timers), B2SV_ADJOINT_RUNS (0: per-gate adjoint sweep), B2SV_LIB (path of an
alternative libb2sv.so for A/B runs inside one box -- boxes differ by up to 15 %).

usage: python benchmarks/tile_experiments.py [--qubits 30] [--layers 4] [--steps 3]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETTINGS = [
    {},                                   # defaults: automatic budget, factored rounds, fused stores
    {"B2SV_TILE_PROF": "1"},              # + phase timers
    {"B2SV_FACTOR": "0"},                 # unfactored dense rounds
    {"B2SV_FUSE_STORE": "0"},             # always the store phase
    {"B2SV_MAX_HEAVY": "4"},              # one round per pass (streaming ceiling of the executor)
    {"B2SV_MAX_HEAVY": "12"},
    {"B2SV_TILE_LOW": "4"},               # 256-byte runs
    {"B2SV_TILE_FLAGS": "1"},             # worker groups in lockstep
    {"B2SV_TILE_FLAGS": "2"},             # worker groups free-running
]


def child(args):
    import ctypes as C

    import numpy as np
    sys.path.insert(0, ROOT)
    from bench import layer_circuit
    from pennylane_lightning_kokkos_b200 import _lib
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

    n = args.qubits
    layers = int(os.environ.get("_layers", args.layers))
    circ = layer_circuit(n, layers, seed=42)
    sv = ops.LightningKokkos_C128(n)
    had = ops.OpsStructKokkos_C128(["Hadamard"] * n, [[] for _ in range(n)], [[w] for w in range(n)],
                                   [False] * n)
    sv.apply_ops(had)
    oplist = ops.OpsStructKokkos_C128([c[0] for c in circ], [c[3] for c in circ], [c[1] for c in circ],
                                      [c[2] for c in circ])
    sv.apply_ops(oplist)
    sv.sync()
    prof = np.zeros(16, dtype=np.uint64)
    _lib.lib.b2sv_debug_tile_prof(prof.ctypes.data_as(C.POINTER(C.c_uint64)))
    sv.reset_stats()
    import subprocess as sp
    mon = sp.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu",
                    "--format=csv,noheader,nounits", "-lms", "50"], stdout=sp.PIPE, text=True)
    time.sleep(0.2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sv.apply_ops(oplist)
    sv.sync()
    dt = (time.perf_counter() - t0) / args.steps
    mon.terminate()
    rows = [[float(x) for x in l.split(",")] for l in mon.stdout.read().splitlines() if l.count(",") == 3]
    rows = rows[4:] or rows
    clk = {"sm_mhz": float(np.median([r[0] for r in rows])), "mem_mhz": float(np.median([r[1] for r in rows])),
           "power_w": float(np.median([r[2] for r in rows])), "temp_c": float(np.max([r[3] for r in rows]))} if rows else {}
    st = sv.stats()
    sweeps = st["sweeps"] / args.steps
    _lib.lib.b2sv_debug_tile_prof(prof.ctypes.data_as(C.POINTER(C.c_uint64)))
    norm = sv.ExpectationValue("Identity", [0], [], np.zeros(0))
    out = {"ms_per_step": dt * 1e3, "sweeps": sweeps, "ms_per_sweep": dt * 1e3 / sweeps,
           "gbs_per_sweep": 2 * 16 * (1 << n) / (dt / sweeps) / 1e9, "norm": norm, "layers": layers, **clk}
    if prof[5]:
        p = [float(x) for x in prof]
        out["prof"] = {"worker_wait_frac": p[0] / (p[0] + p[1]), "producer_wait_frac": p[2] / (p[2] + p[3]),
                       "worker_cycles_per_tile": (p[0] + p[1]) / p[4], "worker_wait_cycles_per_tile": p[0] / p[4],
                       "cta_cycles_per_sweep": p[5] / (148.0 * st["sweeps"]),
                       "last_round_cycles_per_tile": p[6] / p[4], "other_rounds_cycles_per_tile": p[7] / p[4],
                       "worker_prologue_cycles_per_tile": p[10] / p[4], "load_issue_per_tile": p[11] / p[4]}
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    if args.child:
        return child(args)
    for i, s in enumerate(SETTINGS):
        if args.only >= 0 and i != args.only:
            continue
        env = dict(os.environ)
        env.update(s)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--qubits", str(args.qubits),
               "--layers", str(args.layers), "--steps", str(args.steps)]
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=180)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            out = json.loads(res[-1][7:]) if res else {"error": (r.stderr or r.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            out = {"error": "timeout"}
        print(json.dumps({"setting": s, **out}), flush=True)


if __name__ == "__main__":
    main()
