"""CPU check of the identity behind the run differentiation in csrc/adjoint.cpp, in NumPy.

For a run of consecutive single-qubit gates, with W_k the product of the run's gates after gate k,
    <H_lambda_k| G_k |lambda_k>  =  <H_lambda_b| (W_k G_k W_k^dagger)_t |lambda_b>
where t is gate k's wire (gates on other wires cancel), and any 2x2 operator on wire t is a combination
of the four transition sums E_ij(t) = <H_lambda_b| (|i><j|)_t |lambda_b>, which the CUDA kernel gets from
    D = sum_i c_i l_i,  Z_t = sum_i s_t(i) c_i l_i,  X_t = sum_i c_i l_{i^t},  W_t = sum_i s_t(i) c_i l_{i^t}
(c = conj(H_lambda), l = lambda, s_t = +1/-1 by bit t). The Jacobian assembled this way must equal the
oracle's per-gate reverse sweep (reference ADJ.hpp:404-478)."""
import numpy as np

from cases import random_state
from oracle import np_oracle as npo


def transition_sums(h, l, n, wire):
    bit = n - 1 - wire
    idx = np.arange(1 << n)
    s = 1.0 - 2.0 * ((idx >> bit) & 1)
    c = np.conj(h)
    D = np.sum(c * l)
    Z = np.sum(s * c * l)
    X = np.sum(c * l[idx ^ (1 << bit)])
    W = np.sum(s * c * l[idx ^ (1 << bit)])
    return np.array([[0.5 * (D + Z), 0.5 * (X + W)], [0.5 * (X - W), 0.5 * (D - Z)]])  # E_ij


def run_jacobian(final, n, observables, ops, trainable):
    """The whole circuit is one run of single-qubit gates: no gate is undone at all."""
    lam = np.asarray(final, dtype=complex)
    par_index = -1
    entries = []  # (op position, column)
    for pos, (name, wires, inv, params) in enumerate(ops):
        if params:
            par_index += 1
            if par_index in trainable:
                entries.append((pos, trainable.index(par_index)))
    jac = np.zeros((len(observables), len(trainable)))
    for o, ob in enumerate(observables):
        h = npo.apply_obs(lam, n, ob)
        E = {}
        for pos, col in entries:
            name, wires, inv, params = ops[pos]
            w = wires[0]
            acc = np.eye(2, dtype=complex)  # product of the later gates on the same wire, U_b ... U_{k+1}
            for name2, wires2, inv2, params2 in reversed(ops[pos + 1:]):
                if wires2[0] == w:
                    u = npo.gate_matrix(name2, params2)
                    acc = acc @ (u.conj().T if inv2 else u)
            g, scale = npo.generator(name)
            gp = acc @ g @ acc.conj().T
            if w not in E:
                E[w] = transition_sums(h, lam, n, w)
            sign = -1.0 if inv else 1.0
            jac[o, col] = -2.0 * scale * sign * np.imag(np.sum(gp * E[w]))
    return jac


def test_transition_sums_are_matrix_elements():
    n = 5
    h, l = random_state(n, 1), random_state(n, 2)
    for wire in range(n):
        E = transition_sums(h, l, n, wire)
        for i in range(2):
            for j in range(2):
                m = np.zeros((2, 2), dtype=complex)
                m[i, j] = 1.0
                want = np.vdot(h, npo.apply_matrix(l, n, m, [wire]))
                assert abs(E[i, j] - want) < 1e-13


def test_run_differentiation_equals_reverse_sweep():
    n = 6
    rng = np.random.default_rng(3)
    ops = []
    for _ in range(40):
        g = ("RX", "RY", "RZ", "PhaseShift", "Hadamard", "S", "PauliY")[int(rng.integers(7))]
        params = [float(rng.uniform(-2, 2))] if g in ("RX", "RY", "RZ", "PhaseShift") else []
        ops.append((g, [int(rng.integers(n))], bool(rng.integers(2)), params))
    n_par = sum(1 for o in ops if o[3])
    tp = sorted(int(x) for x in rng.choice(n_par, size=(2 * n_par) // 3, replace=False))
    psi0 = random_state(n, 9)
    final = npo.apply_ops(psi0, n, ops)
    observables = [("named", "PauliZ", [0]), ("named", "PauliX", [n - 1]), ("named", "Hadamard", [2])]
    want = npo.adjoint_jacobian(final, n, observables, ops, tp)
    got = run_jacobian(final, n, observables, ops, tp)
    assert np.max(np.abs(want)) > 1e-3
    assert np.max(np.abs(got - want)) < 1e-12
