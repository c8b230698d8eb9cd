"""CPU tests of the PennyLane-free device mirror's host logic (serialisation, error types): the parts of
pennylane_lightning_kokkos_b200/lightning_kokkos.py that need no GPU (reference lightning_kokkos.py /
_serialize.py)."""
import numpy as np
import pytest

from conftest import HAS_GPU


@pytest.fixture(scope="module")
def lk():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos as m
    return m


def test_serialize_ops_follows_the_reference(lk):
    O = lk.Operation
    u = np.array([[0, 1], [1, 0]], dtype=complex)
    tape = lk.QuantumTape([lk.StatePrep(np.array([1, 0], dtype=complex), [0]), O("RX", [0], [0.1]),
                           O("Rot", [1], [0.2, 0.3, 0.4]), O("CNOT", [0, 1]),
                           O("QubitUnitary", [1], matrix=u, adjoint=True), O("CRY", [1, 0], [0.5], adjoint=True)])
    (names, params, wires, inverses, mats), uses_sp = lk._serialize_ops(tape)
    assert uses_sp  # _serialize.py:277-280: state preparation is skipped and reported
    assert names == ["RX", "RZ", "RY", "RZ", "CNOT", "QubitUnitary", "CRY"]  # Rot expanded (_serialize.py:281-282)
    assert params == [[0.1], [0.2], [0.3], [0.4], [], [], [0.5]]
    assert wires == [[0], [1], [1], [1], [0, 1], [1], [1, 0]]
    assert inverses == [False, False, False, False, False, True, True]
    assert [m.size for m in mats] == [0, 0, 0, 0, 0, 4, 0]


def test_serialize_observables_repr_matches_reference_format(lk):
    obs = [lk.NamedObs("PauliZ", [0]), lk.NamedObs("Identity", [1, 2]),
           lk.Tensor([lk.NamedObs("PauliX", [0]), lk.NamedObs("PauliY", [2])]),
           lk.Hamiltonian([0.5, -1.5], [lk.NamedObs("PauliZ", [1]),
                                        lk.Tensor([lk.NamedObs("PauliZ", [0]), lk.NamedObs("Hadamard", [2])])])]
    ser = lk._serialize_observables(lk.QuantumTape([], obs))
    assert repr(ser[0]) == "PauliZ[0]"
    assert repr(ser[1]) == "Identity[1]"  # _serialize.py:75-79: Identity keeps its first wire only
    assert repr(ser[2]) == "PauliX[0] @ PauliY[2]"
    assert repr(ser[3]) == ("Hamiltonian: { 'coeffs' : [0.5, -1.5], 'observables' : "
                            "[PauliZ[1], PauliZ[0] @ Hadamard[2]]}")
    assert ser[3].get_wires() == [0, 1, 2]
    with pytest.raises(lk.QuantumFunctionError):
        lk._serialize_ob(object(), False)


def test_adjoint_operation_check_and_vjp_errors(lk):
    with pytest.raises(lk.QuantumFunctionError):  # lightning_kokkos.py:615-619
        lk.LightningKokkos._check_adjdiff_supported_operations([lk.Operation("CRot", [0, 1], [0.1, 0.2, 0.3])])
    lk.LightningKokkos._check_adjdiff_supported_operations([lk.Operation("Rot", [0], [0.1, 0.2, 0.3])])
    with pytest.raises(TypeError):  # lightning_kokkos.py:177-184
        lk.LightningKokkos(2, c_dtype=np.float32)


@pytest.mark.skipif(HAS_GPU, reason="only meaningful on a box without a CUDA device")
def test_device_needs_a_gpu(lk):
    with pytest.raises(lk.PLException):
        lk.LightningKokkos(2)
