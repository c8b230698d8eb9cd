"""GPU parity at the BASELINE configs' own shapes, against the compiled reference (oracle/_ref), plus
the reference's linear-algebra and Hermitian-in-adjoint literals through the C ABI.

  config 1  20-qubit StronglyEntanglingLayers x4: full state + <Z_i> on all 20 wires
  config 2  26-qubit RX/RY/RZ + CNOT-ring layer: 4096 sampled amplitudes + norm (the reference moves
            ~100 GB for the layer; 28+ qubits would take minutes of host time)
  config 3  20-qubit hardware-efficient ansatz, 7 layers = 420 parameters, 100-term Pauli Hamiltonian:
            the full Jacobian vs the reference's adjointJacobian
  config 4  18-qubit Single/DoubleExcitation circuit + 64-word Pauli sum as CSR: state and CSR expval
Tolerance: 1e-12 relative (complex128), as BASELINE.json's north_star states.
"""
import os
import sys

import numpy as np
import pytest

from cases import layered_circuit, random_pauli_hamiltonian, random_state, sel_circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ops():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    return m


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    return r


def split(circ):
    return ([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ], [c[3] for c in circ])


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def test_config1_at_20_qubits(ops, ref):
    n = 20
    circ = sel_circuit(n, 4, seed=42)
    sv = ops.LightningKokkos_C128(n)
    sv.apply(*split(circ))
    got = np.zeros(1 << n, dtype=np.complex128)
    sv.DeviceToHost(got)
    rsv = ref.RefStateVector(n, np.complex128)
    rsv.apply_ops(circ)
    want = rsv.d2h()
    assert rel_err(got, want) < TOL
    ez = np.array([sv.ExpectationValue("PauliZ", [w], [], np.zeros(0)) for w in range(n)])
    ez_ref = np.array([rsv.expval_named("PauliZ", [w]) for w in range(n)])
    assert np.max(np.abs(ez - ez_ref)) < TOL
    if hasattr(sv, "expval_z_all"):
        assert np.max(np.abs(np.asarray(sv.expval_z_all()) - ez_ref)) < TOL


def test_config2_layer_at_26_qubits_vs_reference(ops, ref):
    n = 26
    layer = layered_circuit(n, 1, seed=42)
    sv = ops.LightningKokkos_C128(n)
    sv.apply(*split(layer))
    rsv = ref.RefStateVector(n, np.complex128)
    rsv.apply_ops(layer)
    rng = np.random.default_rng(11)
    idx = rng.integers(0, 1 << n, size=4096, dtype=np.uint64)
    got = sv.amplitudes(idx)
    want = rsv.amplitudes(idx.astype(np.int64))
    assert rel_err(got, want) < TOL
    assert abs(sv.ExpectationValue("Identity", [0], [], np.zeros(0)) - 1.0) < TOL
    for w in (0, n // 2, n - 1):
        assert abs(sv.ExpectationValue("PauliZ", [w], [], np.zeros(0)) - rsv.expval_named("PauliZ", [w])) < TOL
    assert sv.stats()["sweeps"] <= 8  # fused: a handful of passes for the 104 gates


def test_config3_jacobian_at_20_qubits_vs_reference(ops):
    import configs as cfgs
    r = cfgs.config3_vs_reference(ops, n=20, layers=7, terms=100)
    assert r["params"] == 420
    assert r["ok"], r


def test_config4_excitations_and_csr_vs_reference(ops, ref):
    import scipy.sparse as sp

    n = 18
    rng = np.random.default_rng(42)
    occ, virt = list(range(n // 2)), list(range(n // 2, n))
    circ = []
    for _ in range(50):
        circ.append(("SingleExcitation", [int(rng.choice(occ)), int(rng.choice(virt))], False,
                     [float(rng.uniform(-0.5, 0.5))]))
    for _ in range(100):
        o = [int(x) for x in rng.choice(occ, size=2, replace=False)]
        v = [int(x) for x in rng.choice(virt, size=2, replace=False)]
        circ.append(("DoubleExcitation", o + v, False, [float(rng.uniform(-0.5, 0.5))]))
    ham = random_pauli_hamiltonian(n, 64, seed=7)
    dim = 1 << n
    idx = np.arange(dim, dtype=np.int64)
    mat = sp.csr_matrix((dim, dim), dtype=np.complex128)
    for c, word in ham:
        x = z = ny = 0
        for l, w in word:
            b = 1 << (n - 1 - w)
            if l in ("PauliX", "PauliY"):
                x |= b
            if l in ("PauliZ", "PauliY"):
                z |= b
            ny += l == "PauliY"
        par = np.zeros(dim, dtype=np.int64)
        zz = z
        while zz:
            par ^= (idx >> ((zz & -zz).bit_length() - 1)) & 1
            zz &= zz - 1
        mat = mat + sp.csr_matrix((c * (1j ** ny) * (1 - 2 * par), (idx ^ x, idx)), shape=(dim, dim))
    mat = ((mat + mat.getH()) * 0.5).tocsr()
    mat.sort_indices()
    hf = int("1" * (n // 2) + "0" * (n - n // 2), 2)
    sv = ops.LightningKokkos_C128(n)
    sv.setBasisState(hf)
    sv.apply(*split(circ))
    rsv = ref.RefStateVector(n, np.complex128)
    rsv.set_basis_state(hf)
    rsv.apply_ops(circ)
    got = np.zeros(dim, dtype=np.complex128)
    sv.DeviceToHost(got)
    assert rel_err(got, rsv.d2h()) < TOL
    data, ind, ptr = mat.data, mat.indices.astype(np.uint64), mat.indptr.astype(np.uint64)
    e_ref = rsv.expval_csr(data, ind.astype(np.int64), ptr.astype(np.int64))
    e_cold = sv.ExpectationValue(data, ind, ptr)
    Hs = ops.SparseHamiltonianKokkos_C128(data, ind, ptr, list(range(n)))
    e_res = sv.expval(Hs)
    scale = max(abs(e_ref), 1.0)
    assert abs(e_cold - e_ref) / scale < TOL and abs(e_res - e_ref) / scale < TOL
    # the same Hamiltonian as a Pauli-word sum through the observable classes
    tobs = []
    for _, word in ham:
        fac = [ops.NamedObsKokkos_C128(l, [w]) for l, w in word]
        tobs.append(fac[0] if len(fac) == 1 else ops.TensorProdObsKokkos_C128(fac))
    Hp = ops.HamiltonianKokkos_C128(np.array([c for c, _ in ham]), tobs)
    assert abs(sv.expval(Hp) - e_ref) / scale < 1e-11  # (H + H^dagger)/2 = H for real coefficients


# ---- reference literals: src/tests/Test_LinearAlgebra.cpp:10-91 -----------------------------------
@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-6)])
def test_linear_algebra_literals_through_the_abi(ops, dtype, tol):
    sfx = "C128" if dtype == np.complex128 else "C64"
    SV = getattr(ops, "LightningKokkos_" + sfx)
    # SparseMV (Test_LinearAlgebra.cpp:10-52): y = A x through SparseHamiltonian::applyInPlace
    x = np.array([0, 0.1j, 0.1 + 0.1j, 0.1 + 0.2j, 0.2 + 0.2j, 0.3 + 0.3j, 0.3 + 0.4j, 0.4 + 0.5j])
    want = np.array([0.2 - 0.1j, -0.1 + 0.2j, 0.2 + 0.1j, 0.1 + 0.2j, 0.7 - 0.2j, -0.1 + 0.6j,
                     0.6 + 0.1j, 0.2 + 0.7j])
    indptr = [0, 2, 4, 6, 8, 10, 12, 14, 16]
    indices = [0, 3, 1, 2, 1, 2, 0, 3, 4, 7, 5, 6, 5, 6, 4, 7]
    values = [1, -1j, 1, 1j, -1j, 1, 1j, 1, 1, -1j, 1, 1j, -1j, 1, 1j, 1]
    sv = SV(x.astype(dtype))
    A = getattr(ops, "SparseHamiltonianKokkos_" + sfx)(values, indices, indptr, [0, 1, 2])
    A.apply_in_place(sv)
    got = np.zeros(8, dtype=dtype)
    sv.DeviceToHost(got)
    assert np.max(np.abs(got - want)) < tol
    # axpy (Test_LinearAlgebra.cpp:54-91): v1 += alpha v0
    v0 = np.array([0, 0.1 - 0.1j, 0.1 + 0.1j, 0.2 + 0.1j, 0.2 + 0.2j, 0.3 + 0.3j, 0.4 + 0.3j, 0.5 + 0.4j])
    v1 = np.array([-0.1 + 0.2j, 0.2 - 0.1j, 0.1 + 0.2j, 0.2 + 0.1j, -0.2 + 0.7j, 0.6 - 0.1j, 0.1 + 0.6j,
                   0.7 + 0.2j])
    want = np.array([-0.1 + 0.2j, 0.45 - 0.25j, 0.25 + 0.45j, 0.55 + 0.4j, 0.1 + 1.2j, 1.05 + 0.65j,
                     0.75 + 1.4j, 1.5 + 1.25j])
    s0, s1 = SV(v0.astype(dtype)), SV(v1.astype(dtype))
    s1.axpy(2.0 + 0.5j, s0)
    s1.DeviceToHost(got)
    assert np.max(np.abs(got - want)) < tol
    # inner products (LinearAlgebraKokkos.hpp:155-236) on random vectors, clone and copy
    a, b = random_state(9, 1, dtype), random_state(9, 2, dtype)
    sa, sb = SV(a), SV(b)
    assert abs(sa.inner_product(sb) - np.vdot(a.astype(complex), b.astype(complex))) < tol
    sc = sa.clone()
    assert abs(sc.inner_product(sa) - 1.0) < 10 * tol
    sc.updateData(sb)
    out = np.zeros(1 << 9, dtype=dtype)
    sc.DeviceToHost(out)
    assert np.array_equal(out, b)


# ---- Hermitian and SparseHamiltonian observables inside the adjoint sweep ------------------------
def test_adjoint_hermitian_equals_tensor_literal(ops):
    """src/tests/Test_AdjointDiffKokkos.cpp:457-492: Hermitian(diag(1,-1,-1,1)) on wires {0,1}
    gives the same Jacobian as PauliZ(0) @ PauliZ(1)."""
    param = [-np.pi / 7, np.pi / 5, 2 * np.pi / 3]
    names, wires = ["RX"] * 3, [[0], [1], [2]]
    sv = ops.LightningKokkos_C128(3)
    sv.apply(names, wires, [False] * 3, [[p] for p in param])
    adj = ops.AdjointJacobianKokkos_C128()
    oplist = adj.create_ops_list(names, [np.array([p]) for p in param], wires, [False] * 3,
                                 [np.zeros(0, dtype=complex)] * 3)
    zz = ops.TensorProdObsKokkos_C128([ops.NamedObsKokkos_C128("PauliZ", [0]),
                                       ops.NamedObsKokkos_C128("PauliZ", [1])])
    herm = ops.HermitianObsKokkos_C128(np.diag([1, -1, -1, 1]).astype(complex).ravel(), [0, 1])
    j1 = adj.adjoint_jacobian(sv, [zz], oplist, [0, 2])
    j2 = adj.adjoint_jacobian(sv, [herm], oplist, [0, 2])
    assert np.max(np.abs(j1 - j2)) < 1e-12
    # closed form: <Z0 Z1> = cos(p0) cos(p1)  ->  d/dp0 = -sin(p0) cos(p1), d/dp2 = 0
    assert abs(j1[0, 0] + np.sin(param[0]) * np.cos(param[1])) < 1e-12 and abs(j1[0, 1]) < 1e-12


@pytest.mark.parametrize("n", [5, 12])
def test_adjoint_hermitian_and_sparse_vs_reference(ops, ref, n):
    import scipy.sparse as sp

    circ = layered_circuit(n, 2, seed=13)
    circ += [("CRX", [0, n - 1], False, [0.37]), ("IsingXX", [1, 2], True, [-0.81])]
    names, wires, invs, params = split(circ)
    rng = np.random.default_rng(5)
    m1 = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    m1 = m1 + m1.conj().T
    m2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    m2 = m2 + m2.conj().T
    dim = 1 << n
    S = sp.random(dim, dim, density=min(1.0, 6.0 / dim), random_state=3, dtype=np.float64)
    S = (S + 1j * sp.random(dim, dim, density=min(1.0, 6.0 / dim), random_state=4)).tocsr()
    S = ((S + S.getH()) * 0.5 + sp.identity(dim) * 0.25).tocsr()
    S.sort_indices()
    data, ind, ptr = S.data.astype(complex), S.indices.astype(np.uint64), S.indptr.astype(np.uint64)
    gobs = [ops.HermitianObsKokkos_C128(m1.ravel(), [n - 2]),
            ops.HermitianObsKokkos_C128(m2.ravel(), [0, 3]),
            ops.SparseHamiltonianKokkos_C128(data, ind, ptr, list(range(n))),
            ops.HamiltonianKokkos_C128([0.3, -1.1], [ops.HermitianObsKokkos_C128(m1.ravel(), [1]),
                                                      ops.NamedObsKokkos_C128("PauliX", [2])])]
    robs = [ref.RefObs.hermitian(m1.ravel(), [n - 2]), ref.RefObs.hermitian(m2.ravel(), [0, 3]),
            ref.RefObs.sparse(data, ind.astype(np.int64), ptr.astype(np.int64), list(range(n))),
            ref.RefObs.hamiltonian([0.3, -1.1], [ref.RefObs.hermitian(m1.ravel(), [1]),
                                                 ref.RefObs.named("PauliX", [2])])]
    n_par = sum(1 for p in params if len(p))
    tp = list(range(n_par))
    sv = ops.LightningKokkos_C128(n)
    sv.apply(names, wires, invs, params)
    adj = ops.AdjointJacobianKokkos_C128()
    oplist = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs,
                                 [np.zeros(0, dtype=complex) for _ in names])
    jac = adj.adjoint_jacobian(sv, gobs, oplist, tp)
    rsv = ref.RefStateVector(n, np.complex128)
    rsv.apply_ops(circ)
    jac_ref = rsv.adjoint_jacobian(robs, circ, tp)
    assert rel_err(jac, jac_ref) < TOL


def test_tensor_of_scaled_hamiltonian_factor(ops, ref):
    """A tensor factor that is a one-term Hamiltonian keeps its coefficient (ObservablesKokkos.hpp
    applies every factor in place, coefficient included)."""
    n = 6
    psi = random_state(n, 3)
    sv = ops.LightningKokkos_C128(psi)
    rsv = ref.RefStateVector(n, np.complex128)
    rsv.h2d(psi)
    ob = ops.TensorProdObsKokkos_C128([
        ops.HamiltonianKokkos_C128([2.0], [ops.NamedObsKokkos_C128("PauliZ", [0])]),
        ops.NamedObsKokkos_C128("PauliX", [1])])
    rob = ref.RefObs.tensor([ref.RefObs.hamiltonian([2.0], [ref.RefObs.named("PauliZ", [0])]),
                             ref.RefObs.named("PauliX", [1])])
    assert abs(sv.expval(ob) - rsv.expval_obs(rob)) < TOL
    ham = ops.HamiltonianKokkos_C128([0.7], [ob])
    rham = ref.RefObs.hamiltonian([0.7], [rob])
    assert abs(sv.expval(ham) - rsv.expval_obs(rham)) < TOL
    assert abs(sv.var(ham) - rsv.var_obs(rham)) < 1e-11


def test_adjoint_param_count_check_is_lazy(ops):
    """AdjointDiffKokkos.hpp:444-453: the >1-parameter abort only fires for ops the reverse loop
    visits (the check precedes the break, so the op the loop stops AT is still checked); a Rot
    further in front is never looked at."""
    names = ["Rot", "PauliX", "RX", "RY"]
    wires = [[0], [1], [0], [1]]
    params = [[0.1, 0.2, 0.3], [], [0.4], [0.5]]
    sv = ops.LightningKokkos_C128(2)
    sv.apply(names, wires, [False] * 4, params)
    adj = ops.AdjointJacobianKokkos_C128()
    oplist = adj.create_ops_list(names, [np.array(p) for p in params], wires, [False] * 4,
                                 [np.zeros(0, dtype=complex)] * 4)
    ob = [ops.NamedObsKokkos_C128("PauliZ", [0])]
    jac = adj.adjoint_jacobian(sv, ob, oplist, [1, 2])  # stops before reaching the Rot
    assert jac.shape == (1, 2)
    with pytest.raises(ops.PLException):
        adj.adjoint_jacobian(sv, ob, oplist, [0, 1, 2])  # now the loop reaches the Rot


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 2e-5)])
def test_op_list_handle_replay_matches_reference(ops, ref, dtype, tol):
    """An op-list handle applied again re-uses its schedule and, on small states, a CUDA graph of the
    pass launches; every replay must give what the reference gives for the same gates."""
    n = 14
    sfx = "C128" if dtype == np.complex128 else "C64"
    circ = layered_circuit(n, 3, seed=21) + [("CRX", [0, n - 1], False, [0.4]), ("Toffoli", [3, 1, 7], False, [])]
    names, wires, invs, params = split(circ)
    handle = getattr(ops, "OpsStructKokkos_" + sfx)(names, params, wires, invs)
    sv = getattr(ops, "LightningKokkos_" + sfx)(n)
    rsv = ref.RefStateVector(n, dtype)
    got = np.zeros(1 << n, dtype=dtype)
    for rep in range(3):  # build + capture, replay, replay
        sv.apply_ops(handle)
        rsv.apply_ops(circ)
        sv.DeviceToHost(got)
        assert rel_err(got, rsv.d2h()) < tol * (rep + 1)
    # a second state of another size invalidates the cached plan; a second state of the same size
    # needs its own graph (different buffer)
    sv2 = getattr(ops, "LightningKokkos_" + sfx)(n)
    sv2.apply_ops(handle)
    rsv2 = ref.RefStateVector(n, dtype)
    rsv2.apply_ops(circ)
    sv2.DeviceToHost(got)
    assert rel_err(got, rsv2.d2h()) < tol
    small = getattr(ops, "OpsStructKokkos_" + sfx)(["Hadamard", "CNOT"], [[], []], [[0], [0, 1]], [False, False])
    for nn in (2, 5, 2):
        s3 = getattr(ops, "LightningKokkos_" + sfx)(nn)
        s3.apply_ops(small)
        out = np.zeros(1 << nn, dtype=dtype)
        s3.DeviceToHost(out)
        want = np.zeros(1 << nn, dtype=complex)
        want[0] = want[3 << (nn - 2)] = 2 ** -0.5
        assert np.max(np.abs(out - want)) < tol
    # the <Z> cache must follow graph replays too
    z0 = sv.expval_z_all().copy()
    sv.apply_ops(handle)
    rsv.apply_ops(circ)
    z1 = np.array([rsv.expval_named("PauliZ", [w]) for w in range(n)])
    assert np.max(np.abs(sv.expval_z_all() - z1)) < 10 * tol and np.max(np.abs(z0 - z1)) > 1e-3


def test_bulk_tile_loads_opt_in(ops, ref, monkeypatch):
    """B2SV_BULK=1: passes whose tile starts with >= 8 contiguous index bits use the plain layout and
    bulk async copies (cp.async.bulk); same results as the reference, also when a CTA's three tile
    buffers are re-used many times (24 qubits: ~28 tiles per CTA)."""
    monkeypatch.setenv("B2SV_BULK", "1")
    for n, layers in ((15, 3), (24, 1)):
        circ = layered_circuit(n, layers, seed=31) + [("CRY", [n - 1, 2], False, [0.6]),
                                                      ("MultiRZ", [0, n - 2, 3], False, [0.3])]
        sv = ops.LightningKokkos_C128(n)
        sv.apply(*split(circ))
        rsv = ref.RefStateVector(n, np.complex128)
        rsv.apply_ops(circ)
        idx = np.random.default_rng(2).integers(0, 1 << n, size=4096, dtype=np.uint64)
        assert rel_err(sv.amplitudes(idx), rsv.amplitudes(idx.astype(np.int64))) < TOL
        assert abs(sv.ExpectationValue("Identity", [0], [], np.zeros(0)) - 1.0) < TOL
