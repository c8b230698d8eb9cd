#!/usr/bin/env python
"""One adjoint Jacobian of BASELINE config 3 (for ncu launch lists)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
from cases import random_pauli_hamiltonian
from configs import hea_circuit, split
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops
n = 24
ham = random_pauli_hamiltonian(n, 100, seed=42)
tobs = []
for _, word in ham:
    fac = [ops.NamedObsKokkos_C128(l, [w]) for l, w in word]
    tobs.append(fac[0] if len(fac) == 1 else ops.TensorProdObsKokkos_C128(fac))
H = ops.HamiltonianKokkos_C128(np.array([c for c, _ in ham]), tobs)
circ = hea_circuit(n, 7)
names, wires, invs, params = split(circ)
sv = ops.LightningKokkos_C128(n)
sv.apply(names, wires, invs, params)
adj = ops.AdjointJacobianKokkos_C128()
ol = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs, [np.zeros(0, dtype=complex) for _ in names])
tp = list(range(sum(1 for p in params if len(p))))
for _ in range(2):
    jac = adj.adjoint_jacobian(sv, [H], ol, tp)
print("jac norm", float(np.linalg.norm(jac)))
