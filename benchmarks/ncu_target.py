#!/usr/bin/env python
"""Small driver for ncu captures: BASELINE config-2 layers on an n-qubit c128 state (default 28),
one warm-up step then one profiled step.  usage: python benchmarks/ncu_target.py [qubits] [layers] [c128|c64] [layers|controlled]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import controlled_circuit, layer_circuit  # noqa: E402
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S = "C64" if len(sys.argv) > 3 and sys.argv[3] == "c64" else "C128"
circ = (controlled_circuit if len(sys.argv) > 4 and sys.argv[4] == "controlled" else layer_circuit)(n, layers, seed=42)
sv = getattr(ops, f"LightningKokkos_{S}")(n)
had = getattr(ops, f"OpsStructKokkos_{S}")(["Hadamard"] * n, [[] for _ in range(n)], [[w] for w in range(n)], [False] * n)
sv.apply_ops(had)
ol = getattr(ops, f"OpsStructKokkos_{S}")([c[0] for c in circ], [c[3] for c in circ], [c[1] for c in circ], [c[2] for c in circ])
for _ in range(2):
    sv.apply_ops(ol)
    sv.sync()
print("sweeps", sv.stats()["sweeps"])
