"""CPU tests that pin the oracles (no GPU needed):
  * oracle/_ref (the unmodified reference compiled over the Kokkos stand-in) reproduces the
    literals of the reference's own Catch2 tests;
  * oracle/np_oracle.py (the NumPy restatement) equals oracle/_ref on every gate, generator,
    observable and on the adjoint Jacobian;
  * both equal the committed golden fixtures (tests/golden/*.npz, made by make_golden.py).
"""
import os

import numpy as np
import pytest

from cases import (GATES, GENERATORS, gate_cases, random_circuit, random_pauli_hamiltonian,
                   random_state)
from oracle import np_oracle as npo
from oracle import ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@needs_ref
def test_ref_adjoint_literal():
    """reference src/tests/Test_AdjointDiffKokkos.cpp:213-259"""
    p = [-np.pi / 7, np.pi / 5, 2 * np.pi / 3]
    ops = [("RZ", [0], False, [p[0]]), ("RY", [0], False, [p[1]]), ("RZ", [0], False, [p[2]]),
           ("CNOT", [0, 1], False, []), ("CNOT", [1, 2], False, []), ("RZ", [1], False, [p[0]]),
           ("RY", [1], False, [p[1]]), ("RZ", [1], False, [p[2]])]
    for dt, prec, tol in ((np.complex128, 1, 1e-7), (np.complex64, 0, 1e-5)):
        sv = ref.RefStateVector(3, dt)
        sv.apply_ops(ops)
        obs = ref.RefObs.tensor([ref.RefObs.named("PauliX", [i], prec) for i in range(3)], prec)
        jac = sv.adjoint_jacobian([obs], ops, range(6))
        want = [0.0, -0.674214427, 0.275139672, 0.275139672, -0.0129093062, 0.323846156]
        np.testing.assert_allclose(jac[0], want, atol=tol)


@needs_ref
def test_ref_expval_var_probs_literals():
    """Test_StateVectorKokkos_Expval.cpp:417-491, _Var.cpp:124-144, _Measure.cpp:21-46"""
    init = np.array([0.0, 0.1j, 0.1 + 0.1j, 0.1 + 0.2j, 0.2 + 0.2j, 0.3 + 0.3j, 0.3 + 0.4j, 0.4 + 0.5j])
    sv = ref.RefStateVector(3)
    sv.h2d(init)
    X0, Z1 = ref.RefObs.named("PauliX", [0]), ref.RefObs.named("PauliZ", [1])
    ham = ref.RefObs.hamiltonian([0.3, 0.5], [X0, Z1])
    assert sv.expval_obs(ham) == pytest.approx(-0.086, rel=1e-6)
    assert sv.var_obs(ham) == pytest.approx(0.224604, rel=1e-6)
    assert sv.expval_obs(ref.RefObs.tensor([X0, Z1])) == pytest.approx(-0.36, rel=1e-6)
    index_ptr = [0, 2, 4, 6, 8, 10, 12, 14, 16]
    indices = [0, 3, 1, 2, 1, 2, 0, 3, 4, 7, 5, 6, 5, 6, 4, 7]
    p = 3.1415
    values = [p, -1j * p, p, 1j * p, -1j * p, p, 1j * p, p, p, -1j * p, p, 1j * p, -1j * p, p, 1j * p, p]
    assert sv.expval_csr(values, indices, index_ptr) == pytest.approx(3.1415, rel=1e-7)
    sv2 = ref.RefStateVector(3)
    ph = 0.7
    for q in range(3):
        sv2.apply("RX", [q], False, [ph])
        sv2.apply("RY", [q], False, [ph])
        ph -= 0.2
    np.testing.assert_allclose(sv2.probs([2, 0]), [0.75788676, 0.19844714, 0.03460502, 0.00906107],
                               atol=1e-7)
    np.testing.assert_allclose(sv2.probs([1, 2]), [0.84642778, 0.0386478, 0.10990612, 0.0050183],
                               atol=1e-7)
    # np_oracle on the same literals
    st = sv2.d2h()
    np.testing.assert_allclose(npo.probs(st, 3, [2, 0]), sv2.probs([2, 0]), atol=1e-14)
    np.testing.assert_allclose(npo.probs(st, 3, [1, 2, 0]), sv2.probs([1, 2, 0]), atol=1e-14)
    hnp = ("hamiltonian", [0.3, 0.5], [("named", "PauliX", [0]), ("named", "PauliZ", [1])])
    assert npo.expval(init, 3, hnp) == pytest.approx(-0.086, rel=1e-6)
    assert npo.var(init, 3, hnp) == pytest.approx(0.224604, rel=1e-6)


@needs_ref
@pytest.mark.parametrize("n", [4, 6])
def test_np_oracle_equals_ref_on_every_gate(n):
    st = random_state(n, 5)
    sv = ref.RefStateVector(n)
    worst = 0.0
    for name, wires, inv, params in gate_cases(n, seed=1, per_gate=4):
        sv.h2d(st)
        sv.apply(name, wires, inv, params)
        got = npo.apply_gate(st, n, name, wires, inv, params)
        worst = max(worst, np.max(np.abs(got - sv.d2h())))
    assert worst < 1e-14


@needs_ref
def test_np_oracle_equals_ref_on_generators():
    n = 6
    rng = np.random.default_rng(2)
    st = random_state(n, 6)
    sv = ref.RefStateVector(n)
    for name in GENERATORS:
        nw = GATES[name][0] or 3
        wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
        sv.h2d(st)
        s_ref = sv.apply_generator(name, wires)
        got, s = npo.apply_generator(st, n, name, wires)
        assert s == s_ref
        assert np.max(np.abs(got - sv.d2h())) < 1e-14, name


def test_generators_match_gate_derivatives():
    """Property of Test_StateVectorKokkos_Generator.cpp:17-132: s*G*psi*i == dU/dtheta psi at 0
    shift, checked here by central finite differences of the gate itself."""
    n = 5
    rng = np.random.default_rng(3)
    st = random_state(n, 7)
    eps = 1e-5
    for name in GENERATORS:
        nw = GATES[name][0] or 3
        wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
        theta = 0.37
        up = npo.apply_gate(st, n, name, wires, False, [theta + eps])
        dn = npo.apply_gate(st, n, name, wires, False, [theta - eps])
        fd = (up - dn) / (2 * eps)
        g, s = npo.apply_generator(npo.apply_gate(st, n, name, wires, False, [theta]), n, name, wires)
        assert np.max(np.abs(fd - 1j * s * g)) < 1e-8, name


@needs_ref
def test_np_oracle_adjoint_equals_ref():
    n = 5
    rng = np.random.default_rng(4)
    par = [g for g, (nw, npar) in GATES.items() if npar == 1]
    circ = []
    for g in par:
        nw = GATES[g][0] or 3
        wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
        circ.append((g, wires, bool(rng.integers(2)), [float(rng.uniform(-1, 1))]))
        circ.append(("CNOT", [int(x) for x in rng.choice(n, size=2, replace=False)], False, []))
    terms = random_pauli_hamiltonian(n, 5, seed=1)
    robs, nobs = [], []
    for _, word in terms:
        rf = [ref.RefObs.named(nm, [w]) for nm, w in word]
        robs.append(rf[0] if len(rf) == 1 else ref.RefObs.tensor(rf))
        nf = [("named", nm, [w]) for nm, w in word]
        nobs.append(nf[0] if len(nf) == 1 else ("tensor", nf))
    coeffs = [c for c, _ in terms]
    sv = ref.RefStateVector(n)
    sv.apply_ops(circ)
    n_par = sum(1 for c in circ if c[3])
    tp = sorted(int(x) for x in rng.choice(n_par, size=n_par - 4, replace=False))
    want = sv.adjoint_jacobian([ref.RefObs.hamiltonian(coeffs, robs), robs[1]], circ, tp)
    got = npo.adjoint_jacobian(sv.d2h(), n, [("hamiltonian", coeffs, nobs), nobs[1]], circ, tp)
    assert np.max(np.abs(got - want)) < 1e-13


def test_np_oracle_adjoint_vs_finite_differences():
    n = 4
    circ = random_circuit(n, 30, 9, names=[g for g, (nw, k) in GATES.items() if k <= 1])
    ob = ("hamiltonian", [0.4, -0.7], [("named", "PauliZ", [0]),
                                       ("tensor", [("named", "PauliX", [1]), ("named", "PauliY", [3])])])
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[0] = 1

    def energy(c):
        return npo.expval(npo.apply_ops(psi0, n, c), n, ob)

    par_idx = [i for i, c in enumerate(circ) if c[3]]
    jac = npo.adjoint_jacobian(npo.apply_ops(psi0, n, circ), n, [ob], circ, range(len(par_idx)))
    eps = 1e-6
    for col, i in enumerate(par_idx):
        up = list(circ)
        dn = list(circ)
        up[i] = (circ[i][0], circ[i][1], circ[i][2], [circ[i][3][0] + eps])
        dn[i] = (circ[i][0], circ[i][1], circ[i][2], [circ[i][3][0] - eps])
        fd = (energy(up) - energy(dn)) / (2 * eps)
        assert abs(fd - jac[0, col]) < 1e-7


def test_golden_fixtures_match_np_oracle():
    """tests/golden/*.npz were generated from the compiled reference by tests/golden/make_golden.py."""
    path = os.path.join(GOLD, "gates_n5.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures not generated")
    z = np.load(path, allow_pickle=True)
    st = z["state"]
    for i, case in enumerate(z["cases"]):
        name, wires, inv, params = case
        got = npo.apply_gate(st, 5, name, list(wires), bool(inv), list(params))
        assert np.max(np.abs(got - z["out"][i])) < 1e-14, case
    g = np.load(os.path.join(GOLD, "adjoint_n6.npz"), allow_pickle=True)
    circ = [(c[0], list(c[1]), bool(c[2]), list(c[3])) for c in g["circ"]]
    psi = npo.apply_ops(np.eye(1, 64, 0, dtype=complex).ravel(), 6, circ)
    assert np.max(np.abs(psi - g["state"])) < 1e-13
    nobs = []
    for _, word in g["terms"]:
        nf = [("named", nm, [int(w)]) for nm, w in word]
        nobs.append(nf[0] if len(nf) == 1 else ("tensor", nf))
    ham = ("hamiltonian", [float(c) for c, _ in g["terms"]], nobs)
    jac = npo.adjoint_jacobian(psi, 6, [ham], circ, [int(t) for t in g["tp"]])
    assert np.max(np.abs(jac - g["jac"])) < 1e-13
    assert npo.expval(psi, 6, ham) == pytest.approx(float(g["expval"]), abs=1e-13)


def _ref_param_literals():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_param_literals.json")
    with open(path) as f:
        return json.load(f)["cases"]


def test_reference_param_gate_literals_pin_both_oracles():
    """The literal in/out state vectors of the reference's own parametric-gate tests
    (src/tests/Test_StateVectorKokkos_Param.cpp: IsingXY :24-62, RX :92-150, IsingXX/YY/ZZ :508-855,
    MultiRZ :856-953, SingleExcitation[Minus/Plus] :960-1110, DoubleExcitation[Minus/Plus] :1127-1320)
    and fixed-gate tests (Test_StateVectorKokkos_NonParam.cpp: PauliY/PauliZ/S/T :119-378, SWAP :428-625,
    CZ :626-822, Toffoli :823-931, CSWAP :1011-1113), extracted by tests/golden/make_ref_literals.py,
    pin the compiled reference over the Kokkos stand-in AND the NumPy restatement."""
    cases = _ref_param_literals()
    assert len(cases) >= 73
    assert {"IsingXX", "IsingYY", "IsingZZ", "MultiRZ", "RX", "SWAP", "CZ", "Toffoli", "CSWAP", "PauliY", "PauliZ", "S", "T"} <= {c["gate"] for c in cases}
    for c in cases:
        ini = np.array([complex(a, b) for a, b in c["ini"]])
        want = np.array([complex(a, b) for a, b in c["expected"]])
        n = int(np.log2(ini.size))
        got_np = npo.apply_gate(ini, n, c["gate"], c["wires"], c["inverse"], c["params"])
        assert np.max(np.abs(got_np - want)) < 2e-6, (c["gate"], c["ref_line"])  # the literals carry ~7 digits
        if ref.available():
            sv = ref.RefStateVector(n)
            sv.h2d(ini)
            sv.apply(c["gate"], c["wires"], c["inverse"], c["params"])
            assert np.max(np.abs(sv.d2h() - want)) < 2e-6, (c["gate"], c["ref_line"])


def _measure_kats():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_measure_kats.json")
    with open(path) as f:
        return json.load(f)["cases"]


def test_reference_measurement_known_answers_pin_both_oracles():
    """The reference's own known-answer measurement tests (Test_StateVectorKokkos_Expval.cpp:19-336: named
    observables by functor call and by NamedObs, 1- and 2-qubit matrices; Test_StateVectorKokkos_Var.cpp:19-122:
    var of NamedObs / HermitianObs / TensorProdObs; extracted by tests/golden/make_ref_measure_kats.py)
    pin the NumPy restatement and the compiled reference."""
    cases = _measure_kats()
    assert len(cases) >= 28
    assert {c["kind"] for c in cases} == {"expval", "var"}
    for c in cases:
        n, o = c["n"], c["obs"]
        ops = [(g, w, inv, par) for g, w, inv, par in c["ops"]]
        psi = np.zeros(1 << n, dtype=complex)
        psi[0] = 1
        psi = npo.apply_ops(psi, n, ops)
        if o["type"] == "named":
            nob = ("named", o["name"], o["wires"])
        elif o["type"] == "hermitian":
            k = len(o["wires"])
            mat = np.array([complex(a, b) for a, b in o["matrix"]]).reshape(1 << k, 1 << k)
            nob = ("hermitian", mat, o["wires"])
        else:
            nob = ("tensor", [("named", f["name"], f["wires"]) for f in o["factors"]])
        got = npo.expval(psi, n, nob) if c["kind"] == "expval" else npo.var(psi, n, nob)
        tol = 2e-6 * max(1.0, abs(c["expected"]))  # Catch2 Approx / literals with 6-10 digits
        assert abs(got - c["expected"]) < tol, (c["ref_file"], c["ref_line"])
        if not ref.available():
            continue
        sv = ref.RefStateVector(n)
        sv.apply_ops(ops)
        if o["type"] == "named":
            rob = ref.RefObs.named(o["name"], o["wires"])
            direct = sv.expval_named(o["name"], o["wires"]) if c["kind"] == "expval" else None
        elif o["type"] == "hermitian":
            rob = ref.RefObs.hermitian(mat, o["wires"])
            direct = sv.expval_matrix(mat, o["wires"]) if c["kind"] == "expval" else None
        else:
            rob = ref.RefObs.tensor([ref.RefObs.named(f["name"], f["wires"]) for f in o["factors"]])
            direct = None
        got_r = sv.expval_obs(rob) if c["kind"] == "expval" else sv.var_obs(rob)
        assert abs(got_r - c["expected"]) < tol, (c["ref_file"], c["ref_line"])
        if direct is not None:
            assert abs(direct - c["expected"]) < tol, (c["ref_file"], c["ref_line"])


def _ref_obs(ob):
    if ob[0] == "named":
        return ref.RefObs.named(ob[1], ob[2])
    if ob[0] == "tensor":
        return ref.RefObs.tensor([_ref_obs(o) for o in ob[1]])
    if ob[0] == "hamiltonian":
        return ref.RefObs.hamiltonian(ob[1], [_ref_obs(o) for o in ob[2]])
    raise KeyError(ob[0])


def test_reference_adjoint_known_answers_pin_both_oracles():
    """Every known-answer case of the reference's Test_AdjointDiffKokkos.cpp:34-455 (tests/adjoint_kats.py)
    through the NumPy restatement and the compiled reference."""
    import adjoint_kats
    cases = adjoint_kats.cases()
    assert len(cases) >= 20
    for c in cases:
        n = c["n"]
        psi = np.zeros(1 << n, dtype=complex)
        psi[0] = 1
        if c.get("init") is not None:
            psi = np.array(c["init"], dtype=complex)
        fin = npo.apply_ops(psi, n, c["ops"])
        adjoint_kats.check(npo.adjoint_jacobian(fin, n, c["obs"], c["ops"], c["tp"]), c)
        if ref.available():
            sv = ref.RefStateVector(n)
            sv.h2d(psi)
            sv.apply_ops(c["ops"])
            adjoint_kats.check(sv.adjoint_jacobian([_ref_obs(o) for o in c["obs"]], c["ops"], c["tp"]), c)
