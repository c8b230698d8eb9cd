"""GPU tests of the PennyLane-free device mirror (pennylane_lightning_kokkos_b200/lightning_kokkos.py,
reference lightning_kokkos.py) against the NumPy oracle: one list-apply per tape, state preparation with
the index table built on the device, batched <Z>, var / probability, adjoint Jacobian and vjp."""
import numpy as np
import pytest

from oracle import np_oracle as npo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lk():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos as m
    return m


def rand_unitary(k, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, _ = np.linalg.qr(a)
    return q


def oracle_state(n, prep, ops):
    psi = np.zeros(1 << n, dtype=complex)
    if prep is None:
        psi[0] = 1
    else:
        st, wires = prep
        for v in range(1 << len(wires)):
            idx = 0
            for j, w in enumerate(wires):
                idx |= ((v >> (len(wires) - 1 - j)) & 1) << (n - 1 - w)
            psi[idx] = st[v]
    for op in ops:
        if op.name in npo_named():
            if op.name == "Rot" or len(op.parameters) <= 3:
                psi = npo.apply_gate(psi, n, op.name, list(op.wires), op.adjoint, list(op.parameters))
        else:
            psi = npo.apply_matrix(psi, n, op.matrix, list(op.wires), op.adjoint)
    return psi


def npo_named():
    from cases import GATES
    return set(GATES) | {"Identity"}


def to_npo(o, lk):
    if isinstance(o, lk.NamedObs):
        return ("named", o.name, list(o.wires))
    if isinstance(o, lk.Hermitian):
        return ("hermitian", np.asarray(o.mat), list(o.wires))
    if isinstance(o, lk.Tensor):
        return ("tensor", [to_npo(x, lk) for x in o.obs])
    if isinstance(o, lk.Hamiltonian):
        return ("hamiltonian", list(o.coeffs), [to_npo(x, lk) for x in o.ops])
    raise TypeError(o)


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 2e-5)])
def test_device_tape_execution(lk, dtype, tol):
    n = 14
    rng = np.random.default_rng(3)
    st = rng.normal(size=4) + 1j * rng.normal(size=4)
    st /= np.linalg.norm(st)
    O = lk.Operation
    ops = [O("Hadamard", [w]) for w in range(n)]
    ops += [O("RX", [3], [0.3]), O("Rot", [5], [0.1, -0.7, 1.3]), O("CNOT", [5, 9]), O("Identity", [2]),
            O("CRY", [13, 0], [0.9], adjoint=True), O("IsingXY", [7, 1], [-0.4]),
            O("QubitUnitary", [6, 2], matrix=rand_unitary(2, 1)),
            O("QubitUnitary", [11, 4, 8], matrix=rand_unitary(3, 2), adjoint=True),
            O("DoubleExcitation", [10, 3, 12, 0], [0.77]), O("MultiRZ", [1, 13, 6], [0.21]),
            O("Toffoli", [2, 12, 5]), O("S", [9], adjoint=True)]
    herm = rand_unitary(1, 5)
    herm = herm + herm.conj().T
    obs = [lk.NamedObs("PauliZ", [w]) for w in range(n)]
    obs += [lk.NamedObs("PauliX", [3]), lk.NamedObs("Hadamard", [6]), lk.Hermitian(herm, [8]),
            lk.Tensor([lk.NamedObs("PauliZ", [0]), lk.NamedObs("PauliY", [13])]),
            lk.Hamiltonian([0.3, -1.2, 0.5], [lk.NamedObs("PauliZ", [4]),
                                              lk.Tensor([lk.NamedObs("PauliX", [2]), lk.NamedObs("PauliZ", [7])]),
                                              lk.NamedObs("PauliY", [11])])]
    dev = lk.LightningKokkos(n, c_dtype=dtype)
    tape = lk.QuantumTape([lk.StatePrep(st, [9, 2])] + ops, obs)
    got = dev.execute(tape)
    sweeps = dev._kokkos_state.stats()["sweeps"]
    psi = oracle_state(n, (st, [9, 2]), ops)
    want = np.array([npo.expval(psi, n, to_npo(o, lk)) for o in obs])
    assert np.max(np.abs(got - want)) < tol
    assert np.max(np.abs(dev.state - psi)) / np.max(np.abs(psi)) < tol
    assert sweeps < len(ops)  # one fused list-apply, not a sweep per gate
    # var, probability
    for o in (obs[3], obs[n], obs[n + 2], obs[n + 3]):
        assert abs(dev.var(o) - npo.var(psi, n, to_npo(o, lk))) < 10 * tol
    assert np.max(np.abs(dev.probability([1, 4, 9]) - npo.probs(psi, n, [1, 4, 9]))) < tol
    with pytest.raises(RuntimeError):
        dev.probability([4, 1])
    # the <Z> cache follows the state
    z3 = dev.expval(lk.NamedObs("PauliZ", [3]))
    dev.apply([O("PauliX", [3])])
    assert abs(dev.expval(lk.NamedObs("PauliZ", [3])) + z3) < tol


def test_device_state_preparation_and_errors(lk):
    n = 5
    dev = lk.LightningKokkos(n)
    dev.apply([lk.BasisState([1, 0, 1], [4, 0, 2])])
    psi = dev.state
    assert abs(psi[0b00101] - 1) < 1e-15 and abs(np.linalg.norm(psi) - 1) < 1e-15
    full = np.arange(1, 33, dtype=complex)
    full /= np.linalg.norm(full)
    dev.apply([lk.StatePrep(full, list(range(n)))])
    assert np.allclose(dev.state, full)
    with pytest.raises(ValueError):
        dev.apply([lk.StatePrep(np.ones(4), [0, 1])])  # not normalised
    with pytest.raises(ValueError):
        dev.apply([lk.BasisState([1, 2], [0, 1])])
    with pytest.raises(lk.DeviceError):
        dev.apply([lk.Operation("PauliX", [0]), lk.BasisState([1], [0])])
    with pytest.raises(TypeError):
        lk.LightningKokkos(3, c_dtype=np.float64)


def test_device_adjoint_jacobian_and_vjp(lk):
    n = 6
    O = lk.Operation
    rng = np.random.default_rng(9)
    ops = []
    for layer in range(2):
        for w in range(n):
            ops.append(O("RX", [w], [float(rng.uniform(0, 6))]))
            ops.append(O("Rot", [w], [float(x) for x in rng.uniform(0, 6, size=3)]))
        for w in range(n):
            ops.append(O("CNOT", [w, (w + 1) % n]))
        ops.append(O("IsingZZ", [0, 3], [float(rng.uniform(0, 6))]))
        ops.append(O("CRZ", [5, 1], [float(rng.uniform(0, 6))], adjoint=True))
    herm = rand_unitary(2, 4)
    herm = herm + herm.conj().T
    obs = [lk.NamedObs("PauliZ", [0]), lk.Tensor([lk.NamedObs("PauliX", [1]), lk.NamedObs("PauliZ", [4])]),
           lk.Hermitian(herm, [2, 5]),
           lk.Hamiltonian([0.4, -0.9], [lk.NamedObs("PauliY", [3]), lk.NamedObs("PauliZ", [2])])]
    n_par = sum(len(o.parameters) for o in ops)
    trainable = sorted(rng.choice(n_par, size=17, replace=False).tolist())
    tape = lk.QuantumTape(ops, obs, trainable)
    dev = lk.LightningKokkos(n)
    jac = dev.adjoint_jacobian(tape)
    # oracle: finite differences of the expectation values (independent of the adjoint bookkeeping)
    def run(shift_idx=None, h=0.0):
        k = 0
        psi = np.zeros(1 << n, dtype=complex)
        psi[0] = 1
        for o in ops:
            p = list(o.parameters)
            for j in range(len(p)):
                if k == shift_idx:
                    p[j] += h
                k += 1
            psi = npo.apply_gate(psi, n, o.name, list(o.wires), o.adjoint, p)
        return np.array([npo.expval(psi, n, to_npo(ob, lk)) for ob in obs])
    h = 1e-5
    fd = np.array([(run(t, h) - run(t, -h)) / (2 * h) for t in trainable]).T
    assert jac.shape == (len(obs), len(trainable))
    assert np.max(np.abs(jac - fd)) < 1e-8
    dy = rng.normal(size=len(obs))
    v = dev.vjp(obs, dy)(tape)
    assert np.max(np.abs(v - dy @ jac)) < 1e-12
    assert np.all(dev.vjp(obs, np.zeros(len(obs)))(tape) == 0)
    with pytest.raises(ValueError):
        dev.vjp(obs, dy[:2])
    with pytest.raises(lk.QuantumFunctionError):
        dev.adjoint_jacobian(lk.QuantumTape([O("CRot", [0, 1], [0.1, 0.2, 0.3])], obs[:1], [0]))
