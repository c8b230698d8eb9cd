"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's algorithms for the hot path.

Independent of both the CUDA engine and the compiled reference (oracle/_ref): gates are applied
as dense matrices by tensor contraction, so it also arbitrates where the reference itself is
wrong (generic >=3-wire matrices on non-ascending wires).
Pinned against the compiled reference and the reference's golden literals by
tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product never does.

Conventions (reference simulator/GateFunctors.hpp:32,96-97): wire w <-> bit n-1-w, i.e. the state
reshaped to [2]*n has axis w = wire w; for a k-wire gate wires[0] is the MSB of the local index.
All citations are reference file:line (GF = simulator/GateFunctors.hpp, SV =
simulator/StateVectorKokkos.hpp, MK = simulator/MeasuresKokkos.hpp, ADJ =
algorithms/AdjointDiffKokkos.hpp, OBS = simulator/ObservablesKokkos.hpp).
"""
from __future__ import annotations

import numpy as np

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
P0 = np.array([[1, 0], [0, 0]], dtype=complex)
P1 = np.array([[0, 0], [0, 1]], dtype=complex)


def _kron(*ms):
    out = np.eye(1, dtype=complex)
    for m in ms:
        out = np.kron(out, m)
    return out


def _rot(phi, theta, omega):  # GF:3151-3163
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[np.exp(-0.5j * (phi + omega)) * c, -np.exp(0.5j * (phi - omega)) * s],
                     [np.exp(-0.5j * (phi - omega)) * s, np.exp(0.5j * (phi + omega)) * c]])


def _ctrl(u):  # control = wires[0] (GF:96-97)
    d = u.shape[0]
    m = np.eye(2 * d, dtype=complex)
    m[d:, d:] = u
    return m


def _pair(dim, i, j, blk, rest_phase=1.0):
    m = np.eye(dim, dtype=complex) * rest_phase
    m[np.ix_([i, j], [i, j])] = blk
    return m


def gate_matrix(name, params=(), nwires=None):
    """Dense matrix of a named gate, local index MSB = wires[0]."""
    p = list(params)
    t = p[0] if p else 0.0
    c, s = np.cos(t / 2), np.sin(t / 2)
    em, ep = np.exp(-0.5j * t), np.exp(0.5j * t)
    if name == "Identity":
        return np.eye(1 << (nwires or 1), dtype=complex)
    one = {
        "PauliX": X, "PauliY": Y, "PauliZ": Z, "Hadamard": H,                      # GF:302-433
        "S": np.diag([1, 1j]), "T": np.diag([1, np.exp(0.25j * np.pi)]),           # GF:436-497
    }
    if name in one:
        return one[name].astype(complex)
    if name == "PhaseShift":
        return np.diag([1, np.exp(1j * t)])                                         # GF:500-530
    if name == "RX":
        return np.array([[c, -1j * s], [-1j * s, c]])                               # GF:533-569
    if name == "RY":
        return np.array([[c, -s], [s, c]], dtype=complex)                           # GF:572-608
    if name == "RZ":
        return np.diag([em, ep])                                                    # GF:611-646
    if name == "Rot":
        return _rot(*p)
    if name == "CNOT":
        return _ctrl(X)                                                             # GF:649-691
    if name == "CY":
        return _ctrl(Y)
    if name == "CZ":
        return _ctrl(Z)
    if name == "SWAP":
        return _pair(4, 1, 2, X)                                                    # GF:851-892
    if name == "ControlledPhaseShift":
        return _ctrl(np.diag([1, np.exp(1j * t)]))
    if name in ("CRX", "CRY", "CRZ"):
        return _ctrl(gate_matrix(name[1:], p))
    if name == "CRot":
        return _ctrl(_rot(*p))
    if name == "IsingXX":                                                           # GF:895-958
        return c * np.eye(4) - 1j * s * _kron(X, X)
    if name == "IsingYY":                                                           # GF:1025-1087
        return c * np.eye(4) - 1j * s * _kron(Y, Y)
    if name == "IsingZZ":                                                           # GF:1090-1151
        return np.diag([em, ep, ep, em])
    if name == "IsingXY":                                                           # GF:961-1022
        return _pair(4, 1, 2, np.array([[c, 1j * s], [1j * s, c]]))
    if name in ("SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus"):
        ph = {"SingleExcitation": 1.0, "SingleExcitationMinus": em, "SingleExcitationPlus": ep}[name]
        return _pair(4, 1, 2, np.array([[c, -s], [s, c]]), ph)                      # GF:1154-1339
    if name in ("DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus"):
        ph = {"DoubleExcitation": 1.0, "DoubleExcitationMinus": em, "DoubleExcitationPlus": ep}[name]
        return _pair(16, 3, 12, np.array([[c, -s], [s, c]]), ph)                    # GF:1343-1765
    if name == "CSWAP":
        return _pair(8, 5, 6, X)                                                    # GF:1995-2060
    if name == "Toffoli":
        return _pair(8, 6, 7, X)                                                    # GF:2063-2129
    if name == "MultiRZ":                                                           # GF:2132-2169
        k = nwires
        par = np.array([bin(i).count("1") & 1 for i in range(1 << k)])
        return np.diag(np.where(par == 0, em, ep))
    raise KeyError(name)


def generator(name, nwires=None):
    """(matrix of G, scale s) with U(theta) = exp(i s theta G) (SV:1275-1580, GF:2172-3132)."""
    XX, YY, ZZ = _kron(X, X), _kron(Y, Y), _kron(Z, Z)
    odd2 = np.diag([0, 1, 1, 0]).astype(complex)
    if name in ("RX", "RY", "RZ"):
        return {"RX": X, "RY": Y, "RZ": Z}[name], -0.5
    if name == "PhaseShift":
        return P1, 1.0
    if name == "ControlledPhaseShift":
        return _kron(P1, P1), 1.0
    if name in ("CRX", "CRY", "CRZ"):
        return _kron(P1, {"X": X, "Y": Y, "Z": Z}[name[2]]), -0.5
    if name == "IsingXX":
        return XX, -0.5
    if name == "IsingYY":
        return YY, -0.5
    if name == "IsingZZ":
        return ZZ, -0.5
    if name == "IsingXY":
        return odd2 @ XX, 0.5                      # swap(10,01), 00 and 11 annihilated
    if name == "SingleExcitation":
        return odd2 @ _kron(Y, X), -0.5            # v01' = -i v10, v10' = i v01, rest annihilated
    if name == "SingleExcitationMinus":
        return _pair(4, 1, 2, np.array([[0, -1j], [1j, 0]])), -0.5
    if name == "SingleExcitationPlus":
        return _pair(4, 1, 2, np.array([[0, -1j], [1j, 0]]), -1.0), -0.5
    if name == "DoubleExcitation":
        m = np.zeros((16, 16), dtype=complex)
        m[3, 12], m[12, 3] = -1j, 1j
        return m, -0.5
    if name == "DoubleExcitationMinus":
        return _pair(16, 3, 12, np.array([[0, -1j], [1j, 0]])), -0.5
    if name == "DoubleExcitationPlus":
        return _pair(16, 3, 12, np.array([[0, 1j], [-1j, 0]])), 0.5
    if name == "MultiRZ":
        par = np.array([bin(i).count("1") & 1 for i in range(1 << nwires)])
        return np.diag(np.where(par == 0, 1.0, -1.0)).astype(complex), -0.5
    raise KeyError(name)


def apply_matrix(state, n, matrix, wires, inverse=False):
    """psi <- M psi on `wires` (GF:15-300 semantics, mathematically exact for any wire order)."""
    k = len(wires)
    m = np.asarray(matrix, dtype=complex).reshape(1 << k, 1 << k)
    if inverse:
        m = m.conj().T
    psi = np.asarray(state, dtype=complex).reshape([2] * n)
    psi = np.moveaxis(psi, list(wires), list(range(k)))
    shp = psi.shape
    psi = (m @ psi.reshape(1 << k, -1)).reshape(shp)
    psi = np.moveaxis(psi, list(range(k)), list(wires))
    return np.ascontiguousarray(psi).reshape(-1)


def apply_gate(state, n, name, wires, inverse=False, params=()):
    if name == "Identity":
        return np.asarray(state, dtype=complex)
    return apply_matrix(state, n, gate_matrix(name, params, len(wires)), wires, inverse)


def apply_ops(state, n, ops):
    for name, wires, inverse, params in ops:
        state = apply_gate(state, n, name, wires, inverse, params)
    return state


def apply_generator(state, n, name, wires):
    g, s = generator(name, len(wires))
    return apply_matrix(state, n, g, wires), s


# ---- observables: nested tuples ("named", name, wires) | ("hermitian", matrix, wires) |
# ("tensor", [obs...]) | ("hamiltonian", coeffs, [obs...]) | ("sparse", scipy_csr)
def apply_obs(state, n, ob):
    kind = ob[0]
    if kind == "named":                                   # OBS:116-118
        return apply_gate(state, n, ob[1], ob[2])
    if kind == "hermitian":                               # OBS:173-179
        return apply_matrix(state, n, ob[1], ob[2])
    if kind == "tensor":                                  # OBS:262-266
        for o in ob[1]:
            state = apply_obs(state, n, o)
        return state
    if kind == "hamiltonian":                             # OBS:360-373
        out = np.zeros(1 << n, dtype=complex)
        for c, o in zip(ob[1], ob[2]):
            out = out + c * apply_obs(state, n, o)
        return out
    if kind == "sparse":                                  # OBS:484-494
        return ob[1] @ state
    raise KeyError(kind)


def expval(state, n, ob):                                 # MK:354-360
    return float(np.real(np.vdot(state, apply_obs(state, n, ob))))


def var(state, n, ob):                                    # MK:368-381
    o = apply_obs(state, n, ob)
    return float(np.real(np.vdot(o, o)) - np.real(np.vdot(state, o)) ** 2)


def probs(state, n, wires=None):                          # MK:389-517
    """Marginal probabilities. For UNSORTED wires the reference's transposition kernel
    (MeasuresFunctors.hpp:162-191, driven by MK:493-509) places output digit j at the sorted digit
    argsort[j] instead of rank[j]; its own literals (src/tests/Test_StateVectorKokkos_Measure.cpp:21-46)
    pin that behaviour, so output digit j reports wire wires[argsort[argsort[j]]]. For sorted wires
    (the only case the reference's Python layer allows, lightning_kokkos.py:488-495) this is the
    identity."""
    p = np.abs(np.asarray(state).reshape([2] * n)) ** 2
    if wires is None or list(wires) == list(range(n)):
        return p.reshape(-1)
    wires = list(wires)
    rest = tuple(a for a in range(n) if a not in wires)
    marg = p.sum(axis=rest) if rest else p          # axes = sorted wires
    order = sorted(wires)
    arg = list(np.argsort(wires, kind="stable"))
    eff = [wires[arg[arg[j]]] for j in range(len(wires))]
    return np.transpose(marg, [order.index(w) for w in eff]).reshape(-1)


def adjoint_jacobian(state, n, observables, ops, trainable):
    """Reverse sweep of ADJ:404-478; `state` is the final state U psi0; returns jac[n_obs, n_tp]."""
    tp = list(trainable)
    if not tp:
        raise ValueError("No trainable parameters provided.")
    lam = np.asarray(state, dtype=complex)
    hs = [apply_obs(lam, n, ob) for ob in observables]
    jac = np.zeros((len(observables), len(tp)))
    num_par = sum(1 for o in ops if len(o[3]) > 0)
    cur, tn, it = num_par - 1, len(tp) - 1, len(tp) - 1
    for name, wires, inverse, params in reversed(ops):
        if len(params) > 1:
            raise ValueError("The operation is not supported using the adjoint differentiation method")
        if name in ("StatePrep", "BasisState"):
            continue
        if it < 0:
            break
        mu = lam
        lam = apply_gate(lam, n, name, wires, not inverse, params)          # ADJ:455
        if len(params) > 0:
            if cur == tp[it]:
                gmu, s = apply_generator(mu, n, name, wires)                # ADJ:459-463
                s *= -1.0 if inverse else 1.0
                for o, h in enumerate(hs):
                    jac[o, tn] = -2.0 * s * np.imag(np.vdot(h, gmu))        # ADJ:197-206
                tn -= 1
                it -= 1
            cur -= 1
        hs = [apply_gate(h, n, name, wires, not inverse, params) for h in hs]  # ADJ:476
    return jac
