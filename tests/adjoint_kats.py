"""The reference's adjoint-Jacobian known-answer tests (src/tests/Test_AdjointDiffKokkos.cpp), restated as data:
each case = qubits, optional initial state, op list (name, wires, inverse, params), observables (nested tuples in
the np_oracle form), trainable parameter indices, expected jac[n_obs][n_tp] (None = not checked by the reference),
and the tolerance the reference states. Used by tests/test_oracle.py (both oracles, CPU) and
tests/test_gpu_parity.py (the engine). The op lists / expectations follow the reference line ranges cited per case.
"""
import math

import numpy as np

PARAM = [-math.pi / 7, math.pi / 5, 2 * math.pi / 3]  # Test_AdjointDiffKokkos.cpp:37 (and every other case)
Z = lambda w: ("named", "PauliZ", [w])
X = lambda w: ("named", "PauliX", [w])


def cases():
    out = []
    for p in PARAM:  # :34-59  Op=RX, Obs=Z
        out.append(dict(ref="34-59", n=1, ops=[("RX", [0], False, [p])], obs=[Z(0)], tp=[0],
                        expected=[[-math.sin(p)]], margin=1e-5))
    for p in PARAM:  # :61-80  Op=RY, Obs=X
        out.append(dict(ref="61-80", n=1, ops=[("RY", [0], False, [p])], obs=[X(0)], tp=[0],
                        expected=[[math.cos(p)]], margin=1e-5))
    # :82-108  Op=RX, Obs=[Z,Z]
    out.append(dict(ref="82-108", n=2, ops=[("RX", [0], False, [PARAM[0]])], obs=[Z(0), Z(1)], tp=[0],
                    expected=[[-math.sin(PARAM[0])], [0.0]], margin=1e-7))
    rx3 = [("RX", [i], False, [PARAM[i]]) for i in range(3)]
    # :110-142  Op=[RX,RX,RX], Obs=[Z,Z,Z] (the reference checks the diagonal)
    out.append(dict(ref="110-142", n=3, ops=rx3, obs=[Z(0), Z(1), Z(2)], tp=[0, 1, 2],
                    expected=[[-math.sin(PARAM[0]), None, None], [None, -math.sin(PARAM[1]), None],
                              [None, None, -math.sin(PARAM[2])]], margin=1e-7))
    # :144-178  ... TParams=[0,2]
    out.append(dict(ref="144-178", n=3, ops=rx3, obs=[Z(0), Z(1), Z(2)], tp=[0, 2],
                    expected=[[-math.sin(PARAM[0]), None], [None, 0.0], [None, -math.sin(PARAM[2])]], margin=1e-7))
    # :180-212  Obs=[ZZZ], values "computed with parameter shift"
    out.append(dict(ref="180-212", n=3, ops=rx3, obs=[("tensor", [Z(0), Z(1), Z(2)])], tp=[0, 1, 2],
                    expected=[[-0.1755096592645253, 0.26478810666384334, -0.6312451595102775]], margin=1e-7))
    # :214-259  Op=Mixed, Obs=[XXX]
    mixed = [("RZ", [0], False, [PARAM[0]]), ("RY", [0], False, [PARAM[1]]), ("RZ", [0], False, [PARAM[2]]),
             ("CNOT", [0, 1], False, []), ("CNOT", [1, 2], False, []),
             ("RZ", [1], False, [PARAM[0]]), ("RY", [1], False, [PARAM[1]]), ("RZ", [1], False, [PARAM[2]])]
    out.append(dict(ref="214-259", n=3, ops=mixed, obs=[("tensor", [X(0), X(1), X(2)])], tp=list(range(6)),
                    expected=[[0.0, -0.674214427, 0.275139672, 0.275139672, -0.0129093062, 0.323846156]], margin=1e-7))
    # :262-316  decomposed Rot on (|0> - |1>)/sqrt2, "computed with PennyLane using default.qubit"
    thetas = np.linspace(-2 * math.pi, 2 * math.pi, 7)
    table = [[0, -9.90819496e-01, 0], [-8.18996553e-01, 1.62526544e-01, 0], [-0.203949, 0.48593716, 0], [0, 1, 0],
             [-2.03948985e-01, 4.85937177e-01, 0], [-8.18996598e-01, 1.62526487e-01, 0], [0, -9.90819511e-01, 0]]
    for th, row in zip(thetas, table):
        out.append(dict(ref="262-316", n=1, init=[2 ** -0.5, -(2 ** -0.5)],
                        ops=[("RZ", [0], False, [float(th)]), ("RY", [0], False, [float(th) ** 3]),
                             ("RZ", [0], False, [math.sqrt(2) * float(th)])],
                        obs=[Z(0)], tp=[0, 1, 2], expected=[row], margin=1e-7))
    # :319-389  Mixed Ops, Obs and TParams
    lp = [0.543, 0.54, 0.1, 0.5, 1.3, -2.3, 0.5, -0.5, 0.5]
    names = ["Hadamard", "RX", "CNOT", "RZ", "RY", "RZ", "RZ", "RY", "RZ", "RZ", "RY", "CNOT"]
    wires = [[0], [0], [0, 1], [0], [0], [0], [0], [0], [0], [0], [1], [0, 1]]
    pars = [[], [lp[0]], [], [lp[1]], [lp[2]], [lp[3]], [lp[4]], [lp[5]], [lp[6]], [lp[7]], [lp[8]], []]
    out.append(dict(ref="319-389", n=2, ops=[(a, b, False, c) for a, b, c in zip(names, wires, pars)],
                    obs=[("tensor", [X(0), Z(1)])], tp=[1, 2, 3],
                    expected=[[-0.71429188, 0.04998561, -0.71904837]], margin=0.0))
    # :392-418  Obs=Ham[Z0+Z1]
    out.append(dict(ref="392-418", n=2, ops=[("RX", [0], False, [PARAM[0]])],
                    obs=[("hamiltonian", [0.3, 0.7], [Z(0), Z(1)])], tp=[0],
                    expected=[[-0.3 * math.sin(PARAM[0])]], margin=1e-7))
    # :420-455  Obs=Ham[Z0+Z1+Z2], TParams=[0,2]
    out.append(dict(ref="420-455", n=3, ops=rx3, obs=[("hamiltonian", [0.47, 0.32, 0.96], [Z(0), Z(1), Z(2)])],
                    tp=[0, 2], expected=[[-0.47 * math.sin(PARAM[0]), -0.96 * math.sin(PARAM[2])]], margin=1e-7))
    return out


def check(jac, case, extra_rel=0.0):
    """Catch2's `expected == Approx(x).margin(m)`: |x - expected| <= max(m, eps * (1 + |x|)), eps = 100 * FLT_EPSILON."""
    eps = 1.1920929e-05 + extra_rel
    jac = np.asarray(jac)
    assert jac.shape == (len(case["obs"]), len(case["tp"])), (case["ref"], jac.shape)
    for i, row in enumerate(case["expected"]):
        for j, want in enumerate(row):
            if want is None:
                continue
            got = float(jac[i][j])
            assert abs(got - want) <= max(case["margin"], eps * (1.0 + abs(got))), (case["ref"], i, j, got, want)
