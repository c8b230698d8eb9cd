#!/usr/bin/env python
"""Developer experiment: does the position of the tile's free bits matter?  Times passes whose 2x2
gates sit on a chosen block of index bits (4 gates per pass), 30-qubit c128 state."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

n = 30
sv = ops.LightningKokkos_C128(n)
sv.apply_ops(ops.OpsStructKokkos_C128(["Hadamard"] * n, [[] for _ in range(n)], [[w] for w in range(n)], [False] * n))
rng = np.random.default_rng(0)
for lo in (0, 5, 8, 12, 16, 20, 23, 26):
    bits = [lo + i for i in range(4)]
    names, params, wires = [], [], []
    for rep in range(10):           # 10 dependent groups of 4 gates -> 10 passes with B2SV_MAX_HEAVY=4
        for b in bits:
            names.append("RX"); params.append([float(rng.uniform(0, 6))]); wires.append([n - 1 - b])
            names.append("RY"); params.append([float(rng.uniform(0, 6))]); wires.append([n - 1 - b])
    ol = ops.OpsStructKokkos_C128(names, params, wires, [False] * len(names))
    sv.apply_ops(ol); sv.sync(); sv.reset_stats()
    t0 = time.perf_counter()
    for _ in range(3):
        sv.apply_ops(ol)
    sv.sync()
    dt = (time.perf_counter() - t0) / 3
    sw = sv.stats()["sweeps"] / 3
    print(json.dumps({"gate_bits": bits, "sweeps": sw, "ms_per_sweep": dt * 1e3 / sw,
                      "gbs": 2 * 16 * (1 << n) / (dt / sw) / 1e9}), flush=True)
