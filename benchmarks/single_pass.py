#!/usr/bin/env python
"""Developer experiment: cost of one-gate passes (what the adjoint sweep is made of) on a 24-qubit
c128 state.  usage: python benchmarks/single_pass.py [qubits]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sv = ops.LightningKokkos_C128(n)
for name, wires, params in [("RX", [3], [0.3]), ("RX", [n - 2], [0.3]), ("CNOT", [2, 7], []), ("CRX", [1, 9], [0.2]),
                            ("IsingXX", [4, 11], [0.1]), ("PauliZ", [5], [])]:
    for _ in range(20):
        getattr(sv, name)(wires, False, params)
    sv.sync()
    t0 = time.perf_counter()
    reps = 300
    for _ in range(reps):
        getattr(sv, name)(wires, False, params)
    sv.sync()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"gate": name, "wires": wires, "us_per_call": dt * 1e6,
                      "gbs": 2 * 16 * (1 << n) / dt / 1e9}), flush=True)
