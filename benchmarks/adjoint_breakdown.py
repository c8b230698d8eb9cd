#!/usr/bin/env python
"""Developer experiment: where the adjoint Jacobian time of BASELINE config 3 goes
(Hamiltonian application vs per-layer reverse sweep)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
from cases import random_pauli_hamiltonian
from configs import hea_circuit, split, timed
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

n = 24
ham = random_pauli_hamiltonian(n, 100, seed=42)
tobs = []
for _, word in ham:
    fac = [ops.NamedObsKokkos_C128(l, [w]) for l, w in word]
    tobs.append(fac[0] if len(fac) == 1 else ops.TensorProdObsKokkos_C128(fac))
H = ops.HamiltonianKokkos_C128(np.array([c for c, _ in ham]), tobs)
Z0 = ops.NamedObsKokkos_C128("PauliZ", [0])
adj = ops.AdjointJacobianKokkos_C128()
for layers in (1, 3, 7):
    circ = hea_circuit(n, layers)
    names, wires, invs, params = split(circ)
    sv = ops.LightningKokkos_C128(n)
    sv.apply(names, wires, invs, params)
    ol = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs, [np.zeros(0, dtype=complex) for _ in names])
    tp = list(range(sum(1 for p in params if len(p))))
    t_h, _ = timed(lambda: sv.expval(H), 3, sv.sync)
    t_z, _ = timed(lambda: sv.expval(Z0), 3, sv.sync)
    t_jh, _ = timed(lambda: adj.adjoint_jacobian(sv, [H], ol, tp), 3, sv.sync)
    t_jz, _ = timed(lambda: adj.adjoint_jacobian(sv, [Z0], ol, tp), 3, sv.sync)
    print(json.dumps({"layers": layers, "params": len(tp), "ms_expval_H100": t_h * 1e3, "ms_expval_Z": t_z * 1e3,
                      "ms_jacobian_H100": t_jh * 1e3, "ms_jacobian_Z": t_jz * 1e3}), flush=True)
