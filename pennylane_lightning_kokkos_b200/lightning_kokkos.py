"""PennyLane-free mirror of the reference's device class ``LightningKokkos``
(reference pennylane_lightning_kokkos/lightning_kokkos.py, cited as LK.py) on top of the b2sv binding.

PennyLane itself is a caller of the hot path and is not part of this repository (nor of the image), so
the tape / operation / observable objects the device consumes are restated here as small plain-Python
records with the attributes the reference device reads (``name``, ``wires``, ``parameters``, ``matrix``,
``obs``, ``coeffs`` ...).  Method names, argument meaning and error behaviour follow the reference, so
a maintainer can diff this file against LK.py; what changes is how the binding is driven:

  * ``apply_kokkos`` hands the whole tape to ONE ``apply(names, wires, adjoints, params)`` call (the
    binding has had that overload all along, Bindings.cpp:233-242) instead of one pybind round trip per
    gate (LK.py:371-402), so the fusion scheduler sees the circuit; matrices ride along in the same
    list.  The reference's sticky ``invert_param`` (LK.py:369-377) is not reproduced.
  * ``expval`` of PauliZ observables is served from one read pass over the state for all wires
    (``expval_batch`` asks for several observables at once; LK.py:554-559 issues one call each).
  * ``expval(Hamiltonian)`` goes through the observable classes (one kernel for a sum of Pauli words)
    instead of a dense 2^k x 2^k matrix or a host-built CSR (LK.py:519-534).
  * ``_apply_state_vector_kokkos`` builds the index table on the device (LK.py:317-327 uses
    itertools.product on the host).
  * Hermitian and SparseHamiltonian observables are accepted by ``adjoint_jacobian`` (the reference's
    Python layer refuses them, LK.py:591-603, although its C++ supports them).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import lightning_kokkos_qubit_ops as _ops
from ._lib import PLException  # noqa: F401


class QuantumFunctionError(Exception):
    """Stand-in for pennylane.QuantumFunctionError (LK.py:576-603,615-619)."""


class DeviceError(Exception):
    """Stand-in for pennylane.DeviceError (LK.py:418-420)."""


# ---- the records the device consumes ----------------------------------------------------------------
@dataclass
class Operation:
    """name, wires, parameters as pennylane.operation.Operation exposes them; adjoint=True plays the
    role of qml.adjoint(op) (LK.py:374-376); matrix is used when `name` is not a kernel of the binding."""
    name: str
    wires: Sequence[int]
    parameters: Sequence[float] = ()
    adjoint: bool = False
    matrix: np.ndarray | None = None

    @property
    def num_params(self):
        return len(self.parameters)


@dataclass
class StatePrep:
    state: np.ndarray
    wires: Sequence[int]
    name: str = "StatePrep"

    @property
    def parameters(self):
        return [np.asarray(self.state)]


@dataclass
class BasisState:
    bits: Sequence[int]
    wires: Sequence[int]
    name: str = "BasisState"

    @property
    def parameters(self):
        return [np.asarray(self.bits)]


_PAULI = {
    "Identity": np.eye(2, dtype=complex),
    "PauliX": np.array([[0, 1], [1, 0]], dtype=complex),
    "PauliY": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "PauliZ": np.array([[1, 0], [0, -1]], dtype=complex),
    "Hadamard": np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2),
}


@dataclass
class NamedObs:
    name: str
    wires: Sequence[int]

    def matrix(self):
        return _PAULI[self.name]


@dataclass
class Hermitian:
    mat: np.ndarray
    wires: Sequence[int]
    name: str = "Hermitian"

    def matrix(self):
        return np.asarray(self.mat, dtype=complex)


@dataclass
class Tensor:
    obs: List
    name: str = "Tensor"

    @property
    def wires(self):
        return [w for o in self.obs for w in o.wires]

    def matrix(self):
        m = np.eye(1, dtype=complex)
        for o in self.obs:
            m = np.kron(m, o.matrix())
        return m


@dataclass
class Hamiltonian:
    coeffs: Sequence[float]
    ops: List
    name: str = "Hamiltonian"

    @property
    def wires(self):
        return sorted({w for o in self.ops for w in o.wires})


@dataclass
class SparseHamiltonian:
    csr: object  # scipy.sparse CSR over all device wires
    wires: Sequence[int]
    name: str = "SparseHamiltonian"


@dataclass
class QuantumTape:
    operations: List
    observables: List = field(default_factory=list)  # expectation-value measurements
    trainable_params: List[int] = field(default_factory=list)


# ---- serialisation (reference _serialize.py) ----------------------------------------------------------
def _serialize_ob(o, use_csingle):
    sfx = "C64" if use_csingle else "C128"
    if isinstance(o, NamedObs):  # _serialize.py:75-79
        wires = list(o.wires)[:1] if o.name == "Identity" else list(o.wires)
        return getattr(_ops, "NamedObsKokkos_" + sfx)(o.name, wires)
    if isinstance(o, Hermitian):  # _serialize.py:167-177 (refused upstream, supported here)
        return getattr(_ops, "HermitianObsKokkos_" + sfx)(np.asarray(o.mat, dtype=complex).ravel(), list(o.wires))
    if isinstance(o, Tensor):  # _serialize.py:95-99
        return getattr(_ops, "TensorProdObsKokkos_" + sfx)([_serialize_ob(x, use_csingle) for x in o.obs])
    if isinstance(o, Hamiltonian):  # _serialize.py:113-122
        return getattr(_ops, "HamiltonianKokkos_" + sfx)(np.asarray(o.coeffs, dtype=np.float64),
                                                         [_serialize_ob(x, use_csingle) for x in o.ops])
    if isinstance(o, SparseHamiltonian):  # _serialize.py:136-154
        m = o.csr.tocsr()
        m.sort_indices()
        return getattr(_ops, "SparseHamiltonianKokkos_" + sfx)(
            m.data.astype(complex), m.indices.astype(np.uint64), m.indptr.astype(np.uint64), list(o.wires))
    raise QuantumFunctionError(f"observable {type(o).__name__} is not supported")


def _serialize_observables(tape, use_csingle=False):
    return [_serialize_ob(o, use_csingle) for o in tape.observables]


def _serialize_ops(tape, use_csingle=False):
    """(names, params, wires, inverses, matrices), uses_stateprep -- _serialize.py:277-308; a Rot gate
    is split into RZ RY RZ so that every op carries at most one parameter (_serialize.py:289-291)."""
    names, params, wires, inverses, mats = [], [], [], [], []
    uses_stateprep = False
    for op in tape.operations:
        if isinstance(op, (StatePrep, BasisState)):
            uses_stateprep = True
            continue
        if op.name == "Rot" and not op.adjoint:
            phi, theta, omega = [float(p) for p in op.parameters]
            for nm, p in (("RZ", phi), ("RY", theta), ("RZ", omega)):
                names.append(nm), params.append([p]), wires.append(list(op.wires))
                inverses.append(False), mats.append(np.zeros(0, dtype=complex))
            continue
        names.append(op.name)
        wires.append(list(op.wires))
        inverses.append(bool(op.adjoint))
        if op.name in _ops._GATES_1:
            params.append([float(p) for p in op.parameters])
            mats.append(np.zeros(0, dtype=complex))
        else:
            params.append([])
            mats.append(np.asarray(op.matrix, dtype=complex).ravel())
    return (names, params, wires, inverses, mats), uses_stateprep


# ---- the device ---------------------------------------------------------------------------------------
class LightningKokkos:
    """Mirror of LK.py:139-740 (hot-path methods only; the QubitDevice machinery is PennyLane's)."""

    short_name = "lightning.kokkos"

    def __init__(self, wires, *, sync=True, c_dtype=np.complex128, shots=None, kokkos_args=None):
        if c_dtype is np.complex64:  # LK.py:177-184
            self.use_csingle, self.R_DTYPE = True, np.float32
        elif c_dtype is np.complex128:
            self.use_csingle, self.R_DTYPE = False, np.float64
        else:
            raise TypeError(f"Unsupported complex Type: {c_dtype}")
        self.C_DTYPE = c_dtype
        self.num_wires = int(wires)
        self.wires = list(range(self.num_wires))
        self.shots = shots
        cls = _ops.LightningKokkos_C64 if self.use_csingle else _ops.LightningKokkos_C128
        if kokkos_args is None:
            self._kokkos_state = cls(self.num_wires)
        elif isinstance(kokkos_args, _ops.InitializationSettings):
            self._kokkos_state = cls(self.num_wires, kokkos_args)
        else:
            raise TypeError("Argument kokkos_args must be of type InitializationSettings.")
        self._sync = sync

    # -- state management (LK.py:200-235, 252-327)
    def reset(self):
        self._kokkos_state.resetKokkos()

    def syncH2D(self, state_vector):
        self._kokkos_state.HostToDevice(np.ascontiguousarray(state_vector).ravel(order="C"))

    def syncD2H(self, state_vector):
        self._kokkos_state.DeviceToHost(state_vector.ravel(order="C"))

    @property
    def state(self):
        out = np.zeros(2 ** self.num_wires, dtype=self.C_DTYPE)
        self.syncD2H(out)
        return out

    def _create_basis_state_kokkos(self, index):
        self._kokkos_state.setBasisState(int(index))

    def _apply_state_vector_kokkos(self, state, device_wires):
        state = np.asarray(state, dtype=self.C_DTYPE).ravel()
        device_wires = [int(w) for w in device_wires]
        if state.size != 2 ** len(device_wires):
            raise ValueError("State vector must have shape (2**wires,) or (batch_size, 2**wires).")
        if not np.allclose(np.linalg.norm(state), 1.0, atol=1e-10):  # LK.py:308-310
            raise ValueError("Sum of amplitudes-squared does not equal one.")
        if len(device_wires) == self.num_wires and sorted(device_wires) == device_wires:
            self.syncH2D(state)  # LK.py:312-315
            return
        self._kokkos_state.setStateOnWires(device_wires, state)  # index table built on the device

    def _apply_basis_state_kokkos(self, state, wires):
        state = np.asarray(state)
        wires = [int(w) for w in wires]
        if not set(state.tolist()).issubset({0, 1}):  # LK.py:343-347
            raise ValueError("BasisState parameter must consist of 0 or 1 integers.")
        if len(state) != len(wires):
            raise ValueError("BasisState parameter and wires must be of equal length.")
        num = int(np.dot(state, 2 ** (self.num_wires - 1 - np.array(wires))))
        self._create_basis_state_kokkos(num)

    # -- gates (LK.py:365-420)
    def apply_kokkos(self, operations, **kwargs):
        names, wires, adjoints, params, mats = [], [], [], [], []
        for o in operations:
            if str(o.name) == "Identity":
                continue
            names.append(o.name)
            wires.append([int(w) for w in o.wires])
            if o.name in _ops._GATES_1:
                adjoints.append(bool(o.adjoint))
                params.append([float(p) for p in o.parameters])
                mats.append(None)
            else:
                mat = None if o.matrix is None else np.asarray(o.matrix, dtype=complex)
                if mat is None or mat.size == 0:
                    raise Exception("Unsupported operation")  # LK.py:391-392
                # the reference passes qml.matrix(o), already in inverted form, with inverse = False
                adjoints.append(False)
                params.append([])
                mats.append(mat.conj().T.ravel() if o.adjoint else mat.ravel())
        if not names:
            return
        cls = _ops.OpsStructKokkos_C64 if self.use_csingle else _ops.OpsStructKokkos_C128
        self._kokkos_state.apply_ops(cls(names, params, wires, adjoints, mats))

    def apply(self, operations, **kwargs):
        operations = list(operations)
        if operations:
            if isinstance(operations[0], StatePrep):
                self._apply_state_vector_kokkos(np.array(operations[0].parameters[0]).copy(), operations[0].wires)
                operations = operations[1:]
            elif isinstance(operations[0], BasisState):
                self._apply_basis_state_kokkos(operations[0].parameters[0], operations[0].wires)
                operations = operations[1:]
        for operation in operations:
            if isinstance(operation, (BasisState, StatePrep)):
                raise DeviceError(
                    f"Operation {operation.name} cannot be used after other Operations have already "
                    f"been applied on a {self.short_name} device.")
        self.apply_kokkos(operations)

    # -- measurements (LK.py:424-559)
    def generate_samples(self):
        return self._kokkos_state.GenerateSamples(len(self.wires), self.shots).astype(int)

    def var(self, observable, shot_range=None, bin_size=None):
        m = observable.matrix()
        sqr = m.conj().T @ m
        w = [int(x) for x in observable.wires]
        mean = self._kokkos_state.ExpectationValue(w, m.ravel(order="C"))
        squared_mean = self._kokkos_state.ExpectationValue(w, sqr.ravel(order="C"))
        return squared_mean - mean ** 2

    def probability(self, wires=None, shot_range=None, bin_size=None):
        device_wires = list(self.wires if wires is None else wires)
        if len(device_wires) > 1 and not np.all(np.array(device_wires)[:-1] <= np.array(device_wires)[1:]):
            raise RuntimeError(  # LK.py:488-495
                "Lightning does not currently support out-of-order indices for probabilities")
        return self._kokkos_state.probs(device_wires)

    def expval(self, observable, shot_range=None, bin_size=None):
        if isinstance(observable, NamedObs):
            return self._kokkos_state.ExpectationValue(observable.name, [int(w) for w in observable.wires],
                                                       [], np.zeros(0))
        if isinstance(observable, (Hamiltonian, SparseHamiltonian)):
            return self._kokkos_state.expval(_serialize_ob(observable, self.use_csingle))
        return self._kokkos_state.ExpectationValue([int(w) for w in observable.wires],
                                                   observable.matrix().ravel(order="C"))

    def expval_batch(self, observables):
        """Several expectation values at once: every PauliZ comes out of ONE read pass over the state."""
        z = None
        out = []
        for o in observables:
            if isinstance(o, NamedObs) and o.name == "PauliZ":
                if z is None:
                    z = self._kokkos_state.expval_z_all()
                out.append(float(z[int(o.wires[0])]))
            else:
                out.append(self.expval(o))
        return np.array(out)

    def execute(self, tape):
        self.apply(tape.operations)
        return self.expval_batch(tape.observables)

    # -- adjoint differentiation (LK.py:561-727)
    @staticmethod
    def _check_adjdiff_supported_operations(operations):
        for op in operations:
            if isinstance(op, (StatePrep, BasisState)):
                continue
            if op.num_params > 1 and op.name != "Rot":  # LK.py:615-619
                raise QuantumFunctionError(
                    f'The {op.name} operation is not supported using the "adjoint" differentiation method')

    def adjoint_jacobian(self, tape, starting_state=None, use_device_state=False, **kwargs):
        if len(tape.trainable_params) == 0:
            return np.array(0)
        self._check_adjdiff_supported_operations(tape.operations)
        if starting_state is not None:
            self.syncH2D(np.ravel(starting_state, order="C").astype(self.C_DTYPE))
        elif not use_device_state:
            self.reset()
            self.apply(tape.operations)
        adj = _ops.AdjointJacobianKokkos_C64() if self.use_csingle else _ops.AdjointJacobianKokkos_C128()
        obs_serialized = _serialize_observables(tape, self.use_csingle)
        ops_serialized, use_sp = _serialize_ops(tape, self.use_csingle)
        ops_list = adj.create_ops_list(*ops_serialized)
        # trainable_params index the parametrised records of the tape in order; a Rot contributes three
        # one-parameter ops after serialisation, StatePrep / BasisState none (LK.py:650-669)
        tape_param_to_op_param = {}
        k_tape = k_ops = 0
        for op in tape.operations:
            if isinstance(op, (StatePrep, BasisState)):
                k_tape += 1
                continue
            for _ in op.parameters:
                tape_param_to_op_param[k_tape] = k_ops
                k_tape += 1
                k_ops += 1
        all_params = len(tape.trainable_params)
        tp_shift, record = [], []
        for row, tp in enumerate(sorted(tape.trainable_params)):
            if tp in tape_param_to_op_param:
                tp_shift.append(tape_param_to_op_param[tp])
                record.append(row)
        if not tp_shift:
            return np.zeros((len(tape.observables), all_params))
        jac = np.asarray(adj.adjoint_jacobian(self._kokkos_state, obs_serialized, ops_list, tp_shift))
        jac = jac.reshape(-1, len(tp_shift))
        jac_r = np.zeros((jac.shape[0], all_params))
        jac_r[:, record] = jac
        return jac_r

    def vjp(self, observables, dy, starting_state=None, use_device_state=False):
        dy = np.asarray(dy)
        if np.allclose(dy, 0) or not observables:  # LK.py:701-702
            return lambda tape: np.zeros(len(tape.trainable_params))
        if len(dy) != len(observables):  # LK.py:704-707
            raise ValueError("Number of observables in the tape must be the same as the length of dy in the vjp method")
        if np.iscomplexobj(dy):  # LK.py:709-712
            raise ValueError("The vjp method only works with a real-valued dy when the tape is returning an expectation value")
        ham = Hamiltonian([float(x) for x in dy], list(observables))

        def processing_fn(tape):
            if len(tape.trainable_params) == 0:
                return np.array([], dtype=self.C_DTYPE)
            new_tape = QuantumTape(tape.operations, [ham], tape.trainable_params)
            return self.adjoint_jacobian(new_tape, starting_state, use_device_state).ravel()

        return processing_fn
