#!/usr/bin/env python
"""Adjoint Jacobian on a 30-qubit complex128 state (17.2 GB): how many full-state buffers the sweep
holds and how long it takes (2 ansatz layers = 180 parameters, 8-term Pauli Hamiltonian)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import torch  # noqa: E402  (memory query only)
from cases import random_pauli_hamiltonian  # noqa: E402
from configs import hea_circuit, split  # noqa: E402
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops  # noqa: E402

n, layers, terms = int(sys.argv[1]) if len(sys.argv) > 1 else 30, 2, 8
ham = random_pauli_hamiltonian(n, terms, seed=42)
tobs = []
for _, word in ham:
    fac = [ops.NamedObsKokkos_C128(l, [w]) for l, w in word]
    tobs.append(fac[0] if len(fac) == 1 else ops.TensorProdObsKokkos_C128(fac))
H = ops.HamiltonianKokkos_C128(np.array([c for c, _ in ham]), tobs)
circ = hea_circuit(n, layers)
names, wires, invs, params = split(circ)
sv = ops.LightningKokkos_C128(n)
sv.apply(names, wires, invs, params)
adj = ops.AdjointJacobianKokkos_C128()
ol = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs, [np.zeros(0, dtype=complex) for _ in names])
tp = list(range(sum(1 for p in params if len(p))))
free0, total = torch.cuda.mem_get_info()
peak_used = 0
t0 = time.perf_counter()
jac = adj.adjoint_jacobian(sv, [H], ol, tp)
dt = time.perf_counter() - t0
free1, _ = torch.cuda.mem_get_info()
S = 16.0 * (1 << n)
# finite-difference check of three entries through the forward pass
def energy(shift_k, h):
    s2 = ops.LightningKokkos_C128(n)
    p2 = [list(p) for p in params]
    k = 0
    for i, p in enumerate(p2):
        if p:
            if k == shift_k:
                p2[i] = [p[0] + h]
            k += 1
    s2.apply(names, wires, invs, p2)
    return s2.expval(H)
fd_err = 0.0
for k in (0, 57, len(tp) - 1):
    fd = (energy(k, 1e-4) - energy(k, -1e-4)) / 2e-4
    fd_err = max(fd_err, abs(fd - jac[0, k]))
print(json.dumps({"qubits": n, "params": len(tp), "terms": terms, "s_per_jacobian": dt,
                  "state_GB": S / 1e9, "device_memory_in_use_after_GB": (total - free1) / 1e9,
                  "buffers_of_state_size_kept_in_pools": round((free0 - free1) / S, 2),
                  "traffic_GB": sv.last_adjoint_traffic() / 1e9,
                  "hbm_GBps": sv.last_adjoint_traffic() / dt / 1e9,
                  "max_abs_err_vs_central_differences": fd_err}))
