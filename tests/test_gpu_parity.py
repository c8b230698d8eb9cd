"""GPU parity tests: the CUDA engine (through the C ABI) vs the reference oracle
(oracle/_ref = unmodified reference headers over the Kokkos stand-in) on identical inputs.

Tolerances are BASELINE.json's: 1e-12 relative (complex128) / 1e-5 (complex64) on amplitudes,
expectation values and gradients (relative to the infinity norm of the reference result).
"""
import numpy as np
import pytest

from cases import (GATES, GENERATORS, gate_cases, layered_circuit, random_circuit,
                   random_pauli_hamiltonian, random_state, sel_circuit)

pytestmark = pytest.mark.gpu

TOL = {np.complex128: 1e-12, np.complex64: 1e-5}
DTYPES = [np.complex128, np.complex64]


@pytest.fixture(scope="module", params=["ctypes", "pybind11"])
def ops(request):
    """The binding surface twice: the ctypes mirror and the compiled pybind11 module (same names)."""
    if request.param == "ctypes":
        from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    else:
        from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops_pyb as m
    return m


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    return r


def sv_class(ops, dtype):
    return ops.LightningKokkos_C128 if dtype == np.complex128 else ops.LightningKokkos_C64


def suffix(dtype):
    return "C128" if dtype == np.complex128 else "C64"


def to_host(sv, n, dtype):
    out = np.zeros(1 << n, dtype=dtype)
    sv.DeviceToHost(out)
    return out


def rel_err(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / scale


def apply_gpu(sv, op):
    name, wires, inv, params = op
    getattr(sv, name)(wires, inv, params)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 13, 15])
def test_init_and_basis_state(ops, ref, dtype, n):
    sv = sv_class(ops, dtype)(n)
    r = ref.RefStateVector(n, dtype)
    assert sv.numQubits() == n and sv.dataLength() == 1 << n
    np.testing.assert_array_equal(to_host(sv, n, dtype), r.d2h())
    idx = (1 << n) - 1 if n < 3 else 5
    sv.setBasisState(idx)
    r.set_basis_state(idx)
    np.testing.assert_array_equal(to_host(sv, n, dtype), r.d2h())
    sv.resetKokkos()
    r.reset()
    np.testing.assert_array_equal(to_host(sv, n, dtype), r.d2h())


@pytest.mark.parametrize("dtype", DTYPES)
def test_set_state_vector_and_ctor_from_array(ops, ref, dtype):
    n = 4
    idx = [1, 7, 12]
    vals = np.array([0.5 + 0.5j, -0.5j, 0.5], dtype=np.complex128)
    sv = sv_class(ops, dtype)(n)
    sv.setStateVector(idx, vals)
    r = ref.RefStateVector(n, dtype)
    r.set_state_vector(idx, vals)
    np.testing.assert_array_equal(to_host(sv, n, dtype), r.d2h())
    st = random_state(n, 3, dtype)
    sv2 = sv_class(ops, dtype)(st)
    np.testing.assert_array_equal(to_host(sv2, n, dtype), st)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [4, 6, 14])
def test_every_gate_every_pattern(ops, ref, dtype, n):
    """Each named gate x wire pattern x inverse on a random state (reference GateFunctors.hpp)."""
    st = random_state(n, 11 + n, dtype)
    sv = sv_class(ops, dtype)(n)
    r = ref.RefStateVector(n, dtype)
    worst = 0.0
    for case in gate_cases(n, seed=n, per_gate=4 if n < 14 else 3):
        sv.HostToDevice(st)
        r.h2d(st)
        apply_gpu(sv, case)
        r.apply(*case)
        e = rel_err(to_host(sv, n, dtype), r.d2h())
        worst = max(worst, e)
        assert e < TOL[dtype], f"{case}: rel err {e:.3e}"
    assert worst < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_matrix_ops_1q_2q_and_sorted_kq(ops, ref, dtype):
    """Matrix path. For >=3 wires the reference is only right for ascending wires,
    so those are compared with the reference on ascending wires and with NumPy otherwise."""
    n = 6
    rng = np.random.default_rng(5)
    st = random_state(n, 21, dtype)
    sv = sv_class(ops, dtype)(n)
    r = ref.RefStateVector(n, dtype)
    for wires in ([3], [0], [5], [1, 4], [4, 1], [0, 5], [0, 1, 2], [1, 3, 4], [2, 3, 4, 5]):
        k = len(wires)
        m = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _ = np.linalg.qr(m)
        for inv in (False, True):
            sv.HostToDevice(st)
            r.h2d(st)
            sv.apply("QubitUnitary", wires, inv, [], q.ravel())
            r.apply_matrix(q, wires, inv)
            assert rel_err(to_host(sv, n, dtype), r.d2h()) < TOL[dtype], (wires, inv)


@pytest.mark.parametrize("dtype", DTYPES)
def test_matrix_unsorted_wires_vs_numpy(ops, dtype):
    from oracle import np_oracle
    n = 5
    rng = np.random.default_rng(6)
    st = random_state(n, 22, dtype)
    sv = sv_class(ops, dtype)(n)
    for wires in ([2, 1, 0], [4, 0, 2], [3, 1, 4, 0]):
        k = len(wires)
        m = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        q, _ = np.linalg.qr(m)
        sv.HostToDevice(st)
        sv.apply("QubitUnitary", wires, False, [], q.ravel())
        want = np_oracle.apply_matrix(st.astype(np.complex128), n, q, wires)
        assert rel_err(to_host(sv, n, dtype), want) < (1e-12 if dtype == np.complex128 else 2e-6)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,depth,seed", [(3, 40, 1), (8, 200, 2), (13, 150, 3), (16, 300, 4)])
def test_random_circuits_fused_and_unfused(ops, ref, dtype, n, depth, seed):
    """The fusion scheduler (one list call) and the per-gate path must both equal the reference."""
    circ = random_circuit(n, depth, seed)
    st = random_state(n, seed, dtype)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    r.apply_ops(circ)
    want = r.d2h()
    tol = TOL[dtype] * (10 if dtype == np.complex64 else 1)
    sv = sv_class(ops, dtype)(n)
    sv.HostToDevice(st)
    sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
             [c[3] for c in circ])
    assert rel_err(to_host(sv, n, dtype), want) < tol
    fused_sweeps = sv.stats()["sweeps"]
    sv.set_fusion(False)
    sv.reset_stats()
    sv.HostToDevice(st)
    sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
             [c[3] for c in circ])
    assert rel_err(to_host(sv, n, dtype), want) < tol
    assert sv.stats()["sweeps"] >= fused_sweeps
    # per-gate method calls
    sv.HostToDevice(st)
    for c in circ:
        apply_gpu(sv, c)
    assert rel_err(to_host(sv, n, dtype), want) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_config1_sel_circuit_and_expvals(ops, ref, dtype):
    """BASELINE config 1 (20q StronglyEntanglingLayers x4, <Z_i>) at 16 qubits vs the reference."""
    n = 16
    circ = sel_circuit(n, 4)
    r = ref.RefStateVector(n, dtype)
    r.apply_ops(circ)
    sv = sv_class(ops, dtype)(n)
    sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
             [c[3] for c in circ])
    assert rel_err(to_host(sv, n, dtype), r.d2h()) < TOL[dtype] * (10 if dtype == np.complex64 else 1)
    ez = [sv.ExpectationValue("PauliZ", [w], [], np.zeros(0)) for w in range(n)]
    rz = [r.expval_named("PauliZ", [w]) for w in range(n)]
    assert rel_err(ez, rz) < (1e-12 if dtype == np.complex128 else 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_generators_vs_reference(ops, ref, dtype):
    n = 6
    rng = np.random.default_rng(9)
    st = random_state(n, 31, dtype)
    sv = sv_class(ops, dtype)(n)
    r = ref.RefStateVector(n, dtype)
    for name in GENERATORS:
        nw = GATES[name][0] or 3
        for _ in range(3):
            wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
            sv.HostToDevice(st)
            r.h2d(st)
            s1 = sv.applyGenerator(name, wires, False, [])
            s2 = r.apply_generator(name, wires, False)
            assert s1 == s2, name
            assert rel_err(to_host(sv, n, dtype), r.d2h()) < TOL[dtype], (name, wires)
    with pytest.raises(ops.PLException, match="Generator does not exist"):
        sv.applyGenerator("Hadamard", [0], False, [])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 7, 14])
def test_expvals(ops, ref, dtype, n):
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    st = random_state(n, 41, dtype)
    sv = sv_class(ops, dtype)(st)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    for name in ("Identity", "PauliX", "PauliY", "PauliZ", "Hadamard"):
        for w in sorted({0, n // 2, n - 1}):
            got = sv.ExpectationValue(name, [w], [], np.zeros(0))
            want = r.expval_named(name, [w])
            assert abs(got - want) < tol, (name, w)
    rng = np.random.default_rng(3)
    for wires in ([0], [n - 1], [0, n - 1], [n - 1, 0], [1, 0, 2]):
        if max(wires) >= n or len(set(wires)) != len(wires):
            continue
        k = len(wires)
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        h = a + a.conj().T
        got = sv.ExpectationValue(wires, h.ravel())
        if k >= 3 and wires != sorted(wires):
            continue
        want = r.expval_matrix(h, wires)
        assert abs(got - want) < tol * max(1, abs(want)), wires


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_test_literals(ops, dtype):
    """The reference's own expval / var literals (Test_StateVectorKokkos_Expval.cpp:396-491,
    Test_StateVectorKokkos_Var.cpp:124-144) through the observable classes."""
    S = suffix(dtype)
    init = np.array([0.0, 0.1j, 0.1 + 0.1j, 0.1 + 0.2j, 0.2 + 0.2j, 0.3 + 0.3j, 0.3 + 0.4j, 0.4 + 0.5j],
                    dtype=dtype)
    sv = sv_class(ops, dtype)(init)
    X0 = getattr(ops, f"NamedObsKokkos_{S}")("PauliX", [0])
    Z1 = getattr(ops, f"NamedObsKokkos_{S}")("PauliZ", [1])
    ham = getattr(ops, f"HamiltonianKokkos_{S}")([0.3, 0.5], [X0, Z1])
    ten = getattr(ops, f"TensorProdObsKokkos_{S}")([X0, Z1])
    assert sv.expval(ham) == pytest.approx(-0.086, rel=1e-5)
    assert sv.expval(ten) == pytest.approx(-0.36, rel=1e-5)
    assert sv.var(ham) == pytest.approx(0.224604, rel=1e-5)
    # sparse literal (Expval.cpp:417-446)
    index_ptr = [0, 2, 4, 6, 8, 10, 12, 14, 16]
    indices = [0, 3, 1, 2, 1, 2, 0, 3, 4, 7, 5, 6, 5, 6, 4, 7]
    p = 3.1415
    values = [p, -1j * p, p, 1j * p, -1j * p, p, 1j * p, p, p, -1j * p, p, 1j * p, -1j * p, p, 1j * p, p]
    assert sv.ExpectationValue(np.array(values), indices, index_ptr) == pytest.approx(3.1415, rel=1e-6)
    sp = getattr(ops, f"SparseHamiltonianKokkos_{S}")(values, indices, index_ptr, [0, 1, 2])
    assert sv.expval(sp) == pytest.approx(3.1415, rel=1e-6)
    # 3-qubit Hermitian literal (Expval.cpp:394-415): Re = 1.263
    blk = np.array([[0.5, 0.2 + 0.5j], [0.2 - 0.5j, 0.3]])
    rowa = [0.5, 0.2 + 0.5j] + [0.2 - 0.5j, 0.3] * 3
    rowb = [0.2 - 0.5j, 0.3] * 4
    mat = np.array([rowa, rowb] * 4, dtype=np.complex128)
    herm = getattr(ops, f"HermitianObsKokkos_{S}")(mat.ravel(), [0, 1, 2])
    assert sv.expval(herm) == pytest.approx(1.263, rel=1e-5)
    assert sv.ExpectationValue([0, 1, 2], mat.ravel()) == pytest.approx(1.263, rel=1e-5)
    del blk


@pytest.mark.parametrize("dtype", DTYPES)
def test_observables_vs_reference(ops, ref, dtype):
    n = 8
    S = suffix(dtype)
    prec = 1 if dtype == np.complex128 else 0
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    st = random_state(n, 51, dtype)
    sv = sv_class(ops, dtype)(st)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    terms = random_pauli_hamiltonian(n, 12, seed=4)

    def build(mod_named, mod_tensor, mod_ham):
        obs = []
        for _, word in terms:
            fac = [mod_named(nm, [w]) for nm, w in word]
            obs.append(fac[0] if len(fac) == 1 else mod_tensor(fac))
        return mod_ham([c for c, _ in terms], obs), obs

    gh, gobs = build(getattr(ops, f"NamedObsKokkos_{S}"), getattr(ops, f"TensorProdObsKokkos_{S}"),
                     getattr(ops, f"HamiltonianKokkos_{S}"))
    rh, robs = build(lambda nm, w: ref.RefObs.named(nm, w, prec),
                     lambda f: ref.RefObs.tensor(f, prec),
                     lambda c, o: ref.RefObs.hamiltonian(c, o, prec))
    assert abs(sv.expval(gh) - r.expval_obs(rh)) < tol * 10
    assert abs(sv.var(gh) - r.var_obs(rh)) < tol * 10
    for g, o in zip(gobs, robs):
        assert abs(sv.expval(g) - r.expval_obs(o)) < tol
        assert repr(g) == o.name()
    assert repr(gh).split("'observables'")[1] == rh.name().split("'observables'")[1]
    # Hadamard / Hermitian terms take the generic (copy + apply + axpy) path
    Hd = getattr(ops, f"NamedObsKokkos_{S}")("Hadamard", [2])
    a = np.array([[1.0, 0.5 - 0.2j], [0.5 + 0.2j, -0.3]])
    He = getattr(ops, f"HermitianObsKokkos_{S}")(a.ravel(), [5])
    gh2 = getattr(ops, f"HamiltonianKokkos_{S}")([0.7, -0.4, 0.2], [Hd, He, gobs[0]])
    rh2 = ref.RefObs.hamiltonian([0.7, -0.4, 0.2], [ref.RefObs.named("Hadamard", [2], prec),
                                                    ref.RefObs.hermitian(a, [5], prec), robs[0]], prec)
    assert abs(sv.expval(gh2) - r.expval_obs(rh2)) < tol * 10
    assert abs(sv.var(gh2) - r.var_obs(rh2)) < tol * 10
    with pytest.raises(ops.PLException, match="disjoint"):
        getattr(ops, f"TensorProdObsKokkos_{S}")([Hd, getattr(ops, f"NamedObsKokkos_{S}")("PauliX", [2])])


@pytest.mark.parametrize("dtype", DTYPES)
def test_sparse_hamiltonian_vs_reference(ops, ref, dtype):
    import scipy.sparse as sp
    n = 10
    S = suffix(dtype)
    prec = 1 if dtype == np.complex128 else 0
    rng = np.random.default_rng(8)
    dim = 1 << n
    a = sp.random(dim, dim, density=0.01, random_state=7, format="csr", dtype=np.float64)
    b = sp.random(dim, dim, density=0.01, random_state=8, format="csr", dtype=np.float64)
    h = (a + 1j * b)
    h = (h + h.conj().T).tocsr()
    h.sort_indices()
    st = random_state(n, 61, dtype)
    sv = sv_class(ops, dtype)(st)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    tol = 1e-11 if dtype == np.complex128 else 1e-4
    got = sv.ExpectationValue(h.data, h.indices, h.indptr)
    want = r.expval_csr(h.data, h.indices, h.indptr)
    assert abs(got - want) < tol * max(1.0, abs(want))
    gs = getattr(ops, f"SparseHamiltonianKokkos_{S}")(h.data, h.indices, h.indptr, list(range(n)))
    rs = ref.RefObs.sparse(h.data, h.indices, h.indptr, list(range(n)), prec)
    assert abs(sv.expval(gs) - r.expval_obs(rs)) < tol * max(1.0, abs(want))
    assert abs(sv.var(gs) - r.var_obs(rs)) < 50 * tol * max(1.0, abs(r.var_obs(rs)))
    del rng


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 9, 14])
def test_probs(ops, ref, dtype, n):
    # c64: the reference accumulates marginals in float with atomics; north_star tolerance is 1e-5
    tol = 1e-12 if dtype == np.complex128 else 5e-6
    st = random_state(n, 71, dtype)
    sv = sv_class(ops, dtype)(st)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    np.testing.assert_allclose(sv.probs([]), r.probs(), atol=tol)
    np.testing.assert_allclose(sv.probs(list(range(n))), r.probs(), atol=tol)
    rng = np.random.default_rng(2)
    pats = [[0], [n - 1], [0, 1], [1, 0], [2, 0], [0, 1, 2], [2, 1, 0], [1, 2, 0]]
    pats += [[int(x) for x in rng.choice(n, size=min(n, k), replace=False)] for k in (2, 3, 5, 13)]
    for wires in pats:
        if max(wires) >= n:
            continue
        got = sv.probs(wires)
        want = r.probs(wires)
        np.testing.assert_allclose(got, want, atol=tol, err_msg=str(wires))


def test_probs_reference_literals(ops):
    """Test_StateVectorKokkos_Measure.cpp:21-46 (computed with default.qubit)."""
    sv = ops.LightningKokkos_C128(3)
    ph = 0.7
    for q in range(3):
        sv.RX([q], False, [ph])
        sv.RY([q], False, [ph])
        ph -= 0.2
    lit = {
        (0, 1, 2): [0.67078706, 0.03062806, 0.0870997, 0.00397696, 0.17564072, 0.00801973, 0.02280642, 0.00104134],
        (2, 0): [0.75788676, 0.19844714, 0.03460502, 0.00906107],
        (1, 2): [0.84642778, 0.0386478, 0.10990612, 0.0050183],
        (1,): [0.88507558, 0.11492442],
    }
    for wires, want in lit.items():
        np.testing.assert_allclose(sv.probs(list(wires)), want, atol=1e-7)


@pytest.mark.parametrize("dtype", DTYPES)
def test_sampling_distribution(ops, dtype):
    """Reference test: 100 000 shots histogram within 0.05 of the exact probabilities
    (Test_StateVectorKokkos_Param.cpp:1325-1394). The RNG stream itself is unpinned."""
    n = 4
    sv = sv_class(ops, dtype)(n)
    ph = 0.7
    for q in range(n):
        sv.RX([q], False, [ph])
        sv.RY([q], False, [ph])
        ph -= 0.2
    p = sv.probs([])
    shots = 100000
    s = sv.GenerateSamples(n, shots)
    assert s.shape == (shots, n) and s.dtype == np.uint64
    idx = (s * (1 << np.arange(n - 1, -1, -1, dtype=np.uint64))).sum(axis=1)
    hist = np.bincount(idx.astype(np.int64), minlength=1 << n) / shots
    assert np.max(np.abs(hist - p)) < 0.01
    s2 = sv.GenerateSamples(n, shots)  # same seed every call, like the reference (MK.hpp:551)
    np.testing.assert_array_equal(s, s2)
    # a basis state samples itself; big state exercises the two-level search
    big = sv_class(ops, dtype)(15)
    big.setBasisState(12345)
    sb = big.GenerateSamples(15, 64)
    want = [(12345 >> (14 - j)) & 1 for j in range(15)]
    assert (sb == np.array(want, dtype=np.uint64)).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_adjoint_reference_literals(ops, dtype):
    """Test_AdjointDiffKokkos.cpp:213-259: decomposed Rot, <X x X x X>."""
    S = suffix(dtype)
    p = [-np.pi / 7, np.pi / 5, 2 * np.pi / 3]
    names = ["RZ", "RY", "RZ", "CNOT", "CNOT", "RZ", "RY", "RZ"]
    params = [[p[0]], [p[1]], [p[2]], [], [], [p[0]], [p[1]], [p[2]]]
    wires = [[0], [0], [0], [0, 1], [1, 2], [1], [1], [1]]
    sv = sv_class(ops, dtype)(3)
    sv.apply(names, wires, [False] * 8, params)
    N = getattr(ops, f"NamedObsKokkos_{S}")
    obs = getattr(ops, f"TensorProdObsKokkos_{S}")([N("PauliX", [i]) for i in range(3)])
    adj = getattr(ops, f"AdjointJacobianKokkos_{S}")()
    ol = adj.create_ops_list(names, [np.array(x) for x in params], wires, [False] * 8,
                             [np.zeros(0)] * 8)
    jac = adj.adjoint_jacobian(sv, [obs], ol, list(range(6)))
    want = [0.0, -0.674214427, 0.275139672, 0.275139672, -0.0129093062, 0.323846156]
    np.testing.assert_allclose(jac[0], want, atol=1e-7 if dtype == np.complex128 else 1e-5)
    with pytest.raises(ops.PLException, match="No trainable parameters"):
        adj.adjoint_jacobian(sv, [obs], ol, [])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,layers", [(4, 2), (13, 2)])
def test_adjoint_vs_reference(ops, ref, dtype, n, layers):
    """All parametrised gate families, permuted wires, Hamiltonian + tensor + named observables,
    a trainable subset -- against the reference's AdjointJacobianKokkos."""
    S = suffix(dtype)
    prec = 1 if dtype == np.complex128 else 0
    rng = np.random.default_rng(77)
    par_gates = [g for g, (nw, npar) in GATES.items() if npar == 1]
    circ = []
    for _ in range(layers):
        for g in par_gates:
            nw = GATES[g][0] or 3
            if nw > n:
                continue
            wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
            circ.append((g, wires, bool(rng.integers(2)), [float(rng.uniform(-1, 1))]))
            if rng.integers(2):
                a, b = [int(x) for x in rng.choice(n, size=2, replace=False)]
                circ.append(("CNOT", [a, b], False, []))
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    n_par = sum(1 for c in circ if c[3])
    tp = sorted(int(x) for x in rng.choice(n_par, size=max(1, (2 * n_par) // 3), replace=False))

    terms = random_pauli_hamiltonian(n, 6, seed=12)
    N, T = getattr(ops, f"NamedObsKokkos_{S}"), getattr(ops, f"TensorProdObsKokkos_{S}")
    H = getattr(ops, f"HamiltonianKokkos_{S}")
    gobs, robs = [], []
    for _, word in terms:
        f = [N(nm, [w]) for nm, w in word]
        rf = [ref.RefObs.named(nm, [w], prec) for nm, w in word]
        gobs.append(f[0] if len(f) == 1 else T(f))
        robs.append(rf[0] if len(rf) == 1 else ref.RefObs.tensor(rf, prec))
    coeffs = [c for c, _ in terms]
    g_all = [H(coeffs, gobs), gobs[0], N("Hadamard", [0])]
    r_all = [ref.RefObs.hamiltonian(coeffs, robs, prec), robs[0], ref.RefObs.named("Hadamard", [0], prec)]

    sv = sv_class(ops, dtype)(n)
    sv.apply(names, wires, invs, params)
    r = ref.RefStateVector(n, dtype)
    r.apply_ops(circ)
    adj = getattr(ops, f"AdjointJacobianKokkos_{S}")()
    ol = adj.create_ops_list(names, [np.array(x) for x in params], wires, invs,
                             [np.zeros(0)] * len(names))
    jac = adj.adjoint_jacobian(sv, g_all, ol, tp)
    want = r.adjoint_jacobian(r_all, circ, tp)
    assert jac.shape == want.shape == (3, len(tp))
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    assert rel_err(jac, want) < tol
    # vjp (lightning_kokkos.py:689-727): one reverse sweep over the Hamiltonian sum_i dy_i O_i
    dy = np.array([0.7, -1.3, 0.25])
    got_vjp = adj.vjp(sv, g_all, ol, tp, dy)
    assert got_vjp.shape == (len(tp),)
    assert rel_err(got_vjp, dy @ want) < 10 * tol
    assert np.all(adj.vjp(sv, g_all, ol, tp, np.zeros(3)) == 0)
    with pytest.raises(ValueError):
        adj.vjp(sv, g_all, ol, tp, dy[:2])
    with pytest.raises(ValueError):
        adj.vjp(sv, g_all, ol, tp, dy * 1j)


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_param_gate_literals(ops, dtype):
    """In/out state vectors written out in the reference's own tests
    (src/tests/Test_StateVectorKokkos_Param.cpp and _NonParam.cpp, extracted into
    tests/golden/ref_param_literals.json: 73 gate applications over 19 gate kinds)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_param_literals.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 73
    for c in cases:
        ini = np.array([complex(a, b) for a, b in c["ini"]], dtype=dtype)
        want = np.array([complex(a, b) for a, b in c["expected"]])
        n = int(np.log2(ini.size))
        sv = sv_class(ops, dtype)(n)
        sv.HostToDevice(ini)
        getattr(sv, c["gate"])(c["wires"], c["inverse"], c["params"])
        # the literals carry ~7 digits (the reference compares them with Catch2 Approx, 1.2e-5 relative)
        assert np.max(np.abs(to_host(sv, n, dtype) - want)) < 2e-6, c["gate"]


REF_INI_16 = [  # Test_StateVectorKokkos_Generator.cpp:26-43 (the 4-qubit input every generator test uses)
    (0.267462841882, 0.010768564798), (0.228575129706, 0.010564590956), (0.099492749900, 0.260849823392),
    (0.093690204310, 0.189847108173), (0.033390732374, 0.203836830144), (0.226979395737, 0.081852150975),
    (0.031235505729, 0.176933497281), (0.294287602843, 0.145156781198), (0.152742706049, 0.111628061129),
    (0.012553863703, 0.120027860480), (0.237156555364, 0.154658769755), (0.117001120872, 0.228059505033),
    (0.041495873225, 0.065934827444), (0.089653239407, 0.221581340372), (0.217892322429, 0.291261296999),
    (0.292993251871, 0.186570798697)]


def test_generators_finite_difference_property(ops):
    """The reference's generator tests (Test_StateVectorKokkos_Generator.cpp:17-1175): on its 4-qubit input,
    scale * i * G|psi> equals the central difference (U(ep) - U(-ep))|psi> / (2 ep) of the matching gate,
    ep = 1e-3, margin 1e-4 -- here for every generator, not only the ten the reference spells out."""
    dtype = np.complex128
    ini = np.array([complex(a, b) for a, b in REF_INI_16], dtype=dtype)
    n, ep, margin = 4, 1e-3, 1e-4
    rng = np.random.default_rng(12)
    for name in GENERATORS:
        nw = GATES[name][0] or 3
        for wires in ([1, 0, 2, 3][:nw], [int(x) for x in rng.choice(n, size=nw, replace=False)]):
            g = sv_class(ops, dtype)(ini)
            scale = g.applyGenerator(name, wires, False, [])
            up, um = sv_class(ops, dtype)(ini), sv_class(ops, dtype)(ini)
            getattr(up, name)(wires, False, [ep])
            getattr(um, name)(wires, False, [-ep])
            gen, dp = to_host(g, n, dtype), (to_host(up, n, dtype) - to_host(um, n, dtype)) * (0.5 / ep)
            assert np.max(np.abs(-scale * gen.imag - dp.real)) < margin, (name, wires)
            assert np.max(np.abs(scale * gen.real - dp.imag)) < margin, (name, wires)


@pytest.mark.parametrize("dtype", DTYPES)
def test_multi_qubit_op_equals_named_gate(ops, dtype):
    """Test_StateVectorKokkos_NonParam.cpp:932-1010: the Hadamard / CNOT / Toffoli matrices through the generic
    matrix path give the named gate's result (Toffoli: on every basis state)."""
    n = 4
    h = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    cnot = np.eye(4, dtype=complex)[[0, 1, 3, 2]]
    toff = np.eye(8, dtype=complex)[[0, 1, 2, 3, 4, 5, 7, 6]]
    for name, mat, wires in (("Hadamard", h, [0]), ("CNOT", cnot, [0, 1]), ("Toffoli", toff, [0, 1, 2])):
        for b in (range(1 << n) if name == "Toffoli" else [0]):
            a, m = sv_class(ops, dtype)(n), sv_class(ops, dtype)(n)
            a.setBasisState(b)
            m.setBasisState(b)
            getattr(a, name)(wires, False, [])
            m.apply("QubitUnitary", wires, False, [], mat.ravel())
            assert np.max(np.abs(to_host(a, n, dtype) - to_host(m, n, dtype))) < TOL[dtype], (name, b)


@pytest.mark.parametrize("dtype", DTYPES)
def test_set_state_vector_and_basis_state_literals(ops, dtype):
    """Test_StateVectorKokkos_NonParam.cpp:1114-1200: setStateVector(indices, values) scatters values[i] to
    indices[i] (here: swaps neighbours); setBasisState overwrites a populated state with one 1."""
    init = np.array([0.267462849617 + 0.010768564418j, 0.228575125337 + 0.010564590804j,
                     0.099492751062 + 0.260849833488j, 0.093690201640 + 0.189847111702j,
                     0.015641822883 + 0.225092900621j, 0.205574608177 + 0.082808663337j,
                     0.006827173322 + 0.211631480575j, 0.255280800811 + 0.161572331669j], dtype=dtype)
    expected = init.copy()
    expected[0::2], expected[1::2] = init[1::2], init[0::2]
    sv = sv_class(ops, dtype)(3)
    sv.HostToDevice(init)
    sv.setStateVector([0, 2, 4, 6, 1, 3, 5, 7], np.array([init[1], init[3], init[5], init[7],
                                                          init[0], init[2], init[4], init[6]]))
    np.testing.assert_array_equal(to_host(sv, 3, dtype), expected)
    sv.HostToDevice(init)
    sv.setBasisState(3)
    want = np.zeros(8, dtype=dtype)
    want[3] = 1
    np.testing.assert_array_equal(to_host(sv, 3, dtype), want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_adjoint_reference_known_answers(ops, dtype):
    """Every known-answer case of the reference's Test_AdjointDiffKokkos.cpp:34-455 (tests/adjoint_kats.py:
    analytic -sin / cos values, parameter-shift and default.qubit literals, trainable subsets, tensor and
    Hamiltonian observables, a non-basis initial state) through AdjointJacobianKokkos."""
    import adjoint_kats
    S = suffix(dtype)
    N, T, H = (getattr(ops, f"{k}Kokkos_{S}") for k in ("NamedObs", "TensorProdObs", "Hamiltonian"))

    def build(ob):
        if ob[0] == "named":
            return N(ob[1], ob[2])
        if ob[0] == "tensor":
            return T([build(o) for o in ob[1]])
        return H(ob[1], [build(o) for o in ob[2]])

    adj = getattr(ops, f"AdjointJacobianKokkos_{S}")()
    for c in adjoint_kats.cases():
        n = c["n"]
        if c.get("init") is not None:
            sv = sv_class(ops, dtype)(np.array(c["init"], dtype=dtype))
        else:
            sv = sv_class(ops, dtype)(n)
        names, wires = [o[0] for o in c["ops"]], [o[1] for o in c["ops"]]
        invs, params = [o[2] for o in c["ops"]], [o[3] for o in c["ops"]]
        sv.apply(names, wires, invs, params)
        ol = adj.create_ops_list(names, [np.array(x) for x in params], wires, invs, [np.zeros(0)] * len(names))
        jac = adj.adjoint_jacobian(sv, [build(o) for o in c["obs"]], ol, c["tp"])
        adjoint_kats.check(jac, c, extra_rel=0.0 if dtype == np.complex128 else 2e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_measurement_known_answers(ops, dtype):
    """The reference's known-answer measurement tests (Test_StateVectorKokkos_Expval.cpp:19-336,
    Test_StateVectorKokkos_Var.cpp:19-122, extracted into tests/golden/ref_measure_kats.json): circuit from
    |0..0>, one observable, one number -- through the direct calls and the observable classes."""
    import json
    import os
    S = suffix(dtype)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_measure_kats.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 28
    N = getattr(ops, f"NamedObsKokkos_{S}")
    for c in cases:
        n, o = c["n"], c["obs"]
        sv = sv_class(ops, dtype)(n)
        for g, w, inv, par in c["ops"]:
            getattr(sv, g)(w, inv, par)
        tol = (2e-6 if dtype == np.complex128 else 2e-5) * max(1.0, abs(c["expected"]))
        where = (c["ref_file"], c["ref_line"])
        if o["type"] == "named":
            ob = N(o["name"], o["wires"])
            if c["kind"] == "expval":
                assert abs(sv.ExpectationValue(o["name"], o["wires"], [], np.zeros(0)) - c["expected"]) < tol, where
        elif o["type"] == "hermitian":
            mat = np.array([complex(a, b) for a, b in o["matrix"]])
            ob = getattr(ops, f"HermitianObsKokkos_{S}")(mat, o["wires"])
            if c["kind"] == "expval":
                assert abs(sv.ExpectationValue(o["wires"], mat) - c["expected"]) < tol, where
        else:
            ob = getattr(ops, f"TensorProdObsKokkos_{S}")([N(f["name"], f["wires"]) for f in o["factors"]])
        got = sv.expval(ob) if c["kind"] == "expval" else sv.var(ob)
        assert abs(got - c["expected"]) < tol, where


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2])
def test_one_and_two_qubit_states(ops, ref, dtype, n):
    """States far below one tile (the reference's own tests mostly use 1-3 qubits): every gate that
    fits, a random circuit, named / matrix expvals, var, probs, sampling and the adjoint Jacobian."""
    S = suffix(dtype)
    prec = 1 if dtype == np.complex128 else 0
    tol = TOL[dtype]
    st = random_state(n, 90 + n, dtype)
    sv = sv_class(ops, dtype)(n)
    r = ref.RefStateVector(n, dtype)
    for case in gate_cases(n, seed=3, per_gate=2):
        sv.HostToDevice(st)
        r.h2d(st)
        apply_gpu(sv, case)
        r.apply(*case)
        assert rel_err(to_host(sv, n, dtype), r.d2h()) < tol, case
    circ = random_circuit(n, 30, seed=5)
    sv.HostToDevice(st)
    r.h2d(st)
    sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ], [c[3] for c in circ])
    r.apply_ops(circ)
    assert rel_err(to_host(sv, n, dtype), r.d2h()) < 10 * tol
    N = getattr(ops, f"NamedObsKokkos_{S}")
    for name in ("Identity", "PauliX", "PauliY", "PauliZ", "Hadamard"):
        for w in range(n):
            assert abs(sv.ExpectationValue(name, [w], [], np.zeros(0)) - r.expval_named(name, [w])) < 10 * tol
            assert abs(sv.var(N(name, [w])) - r.var_obs(ref.RefObs.named(name, [w], prec))) < 10 * tol
    rng = np.random.default_rng(6)
    for wires in ([0], [n - 1], list(range(n)), list(range(n))[::-1]):
        k = len(wires)
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        h = a + a.conj().T
        want = r.expval_matrix(h, wires)
        assert abs(sv.ExpectationValue(wires, h.ravel()) - want) < 10 * tol * max(1, abs(want)), wires
    ptol = 1e-12 if dtype == np.complex128 else 5e-6
    np.testing.assert_allclose(sv.probs([]), r.probs(), atol=ptol)
    for wires in ([0], [n - 1], list(range(n))[::-1]):
        np.testing.assert_allclose(sv.probs(wires), r.probs(wires), atol=ptol, err_msg=str(wires))
    shots = 20000
    smp = sv.GenerateSamples(n, shots)
    assert smp.shape == (shots, n)
    idx = (smp * (1 << np.arange(n - 1, -1, -1, dtype=np.uint64))).sum(axis=1)
    hist = np.bincount(idx.astype(np.int64), minlength=1 << n) / shots
    assert np.max(np.abs(hist - r.probs())) < 0.02
    # adjoint: RX RY (CNOT) RZ per wire, <Z_0>, <X_{n-1}> and their tensor / sum
    circ = [("RX", [0], False, [0.4]), ("RY", [n - 1], True, [-0.9])]
    if n == 2:
        circ += [("CNOT", [0, 1], False, []), ("IsingXY", [1, 0], False, [0.3])]
    circ += [("RZ", [0], False, [1.1]), ("PhaseShift", [n - 1], False, [0.2])]
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    sv2 = sv_class(ops, dtype)(n)
    sv2.apply(names, wires, invs, params)
    r2 = ref.RefStateVector(n, dtype)
    r2.apply_ops(circ)
    g_all = [N("PauliZ", [0]), N("PauliX", [n - 1]),
             getattr(ops, f"HamiltonianKokkos_{S}")([0.5, -0.25], [N("PauliZ", [0]), N("PauliY", [n - 1])])]
    r_all = [ref.RefObs.named("PauliZ", [0], prec), ref.RefObs.named("PauliX", [n - 1], prec),
             ref.RefObs.hamiltonian([0.5, -0.25], [ref.RefObs.named("PauliZ", [0], prec),
                                                   ref.RefObs.named("PauliY", [n - 1], prec)], prec)]
    adj = getattr(ops, f"AdjointJacobianKokkos_{S}")()
    ol = adj.create_ops_list(names, [np.array(x) for x in params], wires, invs, [np.zeros(0)] * len(names))
    tp = list(range(sum(1 for c in circ if c[3])))
    jac = adj.adjoint_jacobian(sv2, g_all, ol, tp)
    want = r2.adjoint_jacobian(r_all, circ, tp)
    assert rel_err(jac, want) < (1e-12 if dtype == np.complex128 else 2e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_adjoint_runs_equal_per_gate_path(ops, dtype, monkeypatch):
    """Runs of single-qubit gates are differentiated from batched transition sums (adjoint.cpp); the
    result must equal the per-gate sweep (B2SV_ADJOINT_RUNS=0) on a 20-qubit ansatz with inverse
    gates, PhaseShift, fixed single-qubit gates inside the runs and a trainable subset."""
    S = suffix(dtype)
    n = 20
    rng = np.random.default_rng(5)
    circ = []
    for layer in range(3):
        for w in range(n):
            for g in ("RX", "RY", "RZ", "PhaseShift"):
                if rng.integers(4):
                    circ.append((g, [w], bool(rng.integers(2)), [float(rng.uniform(-2, 2))]))
            if rng.integers(3) == 0:
                circ.append((("Hadamard", "S", "PauliY", "T")[int(rng.integers(4))], [w], bool(rng.integers(2)), []))
        for w in range(n):
            circ.append(("CNOT", [w, (w + 1 + layer) % n], False, []))
        circ.append(("CRY", [1, 7], False, [0.3]))
        circ.append(("IsingXX", [3, 12], False, [-0.4]))
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    n_par = sum(1 for c in circ if c[3])
    tp = sorted(int(x) for x in rng.choice(n_par, size=(3 * n_par) // 4, replace=False))
    N, T = getattr(ops, f"NamedObsKokkos_{S}"), getattr(ops, f"TensorProdObsKokkos_{S}")
    H = getattr(ops, f"HamiltonianKokkos_{S}")
    terms = random_pauli_hamiltonian(n, 12, seed=3)
    words = [(lambda f: f[0] if len(f) == 1 else T(f))([N(nm, [w]) for nm, w in word]) for _, word in terms]
    obs = [H([c for c, _ in terms], words), N("PauliZ", [0]), N("PauliX", [n - 1])]
    sv = sv_class(ops, dtype)(n)
    sv.apply(names, wires, invs, params)
    adj = getattr(ops, f"AdjointJacobianKokkos_{S}")()
    ol = adj.create_ops_list(names, [np.array(x) for x in params], wires, invs, [np.zeros(0)] * len(names))
    monkeypatch.setenv("B2SV_ADJOINT_RUNS", "1")
    jac_runs = adj.adjoint_jacobian(sv, obs, ol, tp)
    monkeypatch.setenv("B2SV_ADJOINT_RUNS", "0")
    jac_gate = adj.adjoint_jacobian(sv, obs, ol, tp)
    assert jac_runs.shape == (3, len(tp))
    assert np.max(np.abs(jac_gate)) > 1e-3
    assert rel_err(jac_runs, jac_gate) < (1e-12 if dtype == np.complex128 else 5e-5)


def test_config2_layer_properties_at_scale(ops):
    """BASELINE config 2 shape at 26 qubits (1 GiB state): norm preserved, U^dagger U = identity,
    sampled amplitudes equal the per-gate (unfused) path."""
    n = 26
    circ = layered_circuit(n, 1, seed=42)
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    sv = ops.LightningKokkos_C128(n)
    for w in range(n):
        sv.Hadamard([w], False, [])
    sv.reset_stats()
    sv.apply(names, wires, invs, params)
    fused = sv.stats()["sweeps"]
    assert fused <= 8, f"fusion regressed: {fused} sweeps for one layer"
    assert abs(sv.ExpectationValue("Identity", [0], [], np.zeros(0)) - 1.0) < 1e-12
    probe = ops.LightningKokkos_C128(n)
    for w in range(n):
        probe.Hadamard([w], False, [])
    probe.set_fusion(False)
    probe.apply(names, wires, invs, params)
    ol = ops.OpsStructKokkos_C128(names, params, wires, invs)
    a = np.zeros(1 << n, dtype=np.complex128)
    b = np.zeros(1 << n, dtype=np.complex128)
    sv.DeviceToHost(a)
    probe.DeviceToHost(b)
    idx = np.random.default_rng(0).integers(0, 1 << n, size=4096)
    assert np.max(np.abs(a[idx] - b[idx])) < 1e-12 * np.max(np.abs(b))
    sv.apply_ops(ol, adjoint=True)  # undo the layer
    sv.DeviceToHost(a)
    assert np.max(np.abs(a - 2.0 ** (-n / 2))) < 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("max_heavy", [8, 16])
def test_factored_dense_rounds_vs_reference(ops, ref, dtype, max_heavy, monkeypatch):
    """Layered circuits run as factored dense rounds (D * shear * D form, anti-diagonal-dominant
    gates split into X * (X M)); they must equal the reference and the unfactored executor."""
    n = 14
    circ = layered_circuit(n, 3, seed=5)
    for w in range(n):  # diagonals that vanish (exactly or almost): the X-split path
        circ.append(("RX", [w], False, [np.pi - 1e-3 * w]))
        circ.append(("RY", [(w + 3) % n], False, [np.pi + 1e-9 * w]))
        circ.append(("PauliY", [(w + 5) % n], False, []))
        circ.append(("Hadamard", [(w + 7) % n], False, []))
    st = random_state(n, 77, dtype)
    r = ref.RefStateVector(n, dtype)
    r.h2d(st)
    r.apply_ops(circ)
    want = r.d2h()
    tol = TOL[dtype] * (10 if dtype == np.complex64 else 1)
    monkeypatch.setenv("B2SV_MAX_HEAVY", str(max_heavy))
    for factor in ("1", "0"):
        monkeypatch.setenv("B2SV_FACTOR", factor)
        sv = sv_class(ops, dtype)(n)
        sv.HostToDevice(st)
        sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
                 [c[3] for c in circ])
        assert rel_err(to_host(sv, n, dtype), want) < tol, factor
