"""Mirror of the reference's pybind11 module ``lightning_kokkos_qubit_ops``
(reference pennylane_lightning_kokkos/src/bindings/Bindings.cpp), over the b2sv C ABI.

Same class names, method names, argument order and error type (``PLException``), so the
reference's Python layer (``lightning_kokkos.py`` / ``_serialize.py``) can bind to this module
unchanged.  Per precision: ``LightningKokkos_C64/_C128`` (Bindings.cpp:59-585),
``NamedObsKokkos_*``, ``HermitianObsKokkos_*``, ``TensorProdObsKokkos_*``, ``HamiltonianKokkos_*``,
``SparseHamiltonianKokkos_*`` (:591-736), ``OpsStructKokkos_*`` (:741-762),
``AdjointJacobianKokkos_*`` (:768-821); module functions ``kokkos_start/kokkos_end/
kokkos_config_info/print_configuration`` (:842-852) and ``InitializationSettings`` (:854-967).
"""
from __future__ import annotations

import ctypes as C

from itertools import chain as _itertools_chain

import numpy as np

from . import _lib
from ._lib import PLException, check, lib  # noqa: F401

_GATES_1 = ["Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "T", "CNOT", "SWAP", "CSWAP",
            "Toffoli", "CY", "CZ", "PhaseShift", "ControlledPhaseShift", "RX", "RY", "RZ", "Rot",
            "CRX", "CRY", "CRZ", "CRot", "IsingXX", "IsingXY", "IsingYY", "IsingZZ", "MultiRZ",
            "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus", "DoubleExcitation",
            "DoubleExcitationMinus", "DoubleExcitationPlus"]


def _chain(lists):
    return _itertools_chain.from_iterable(lists)


def _i64(a):
    arr = np.ascontiguousarray(a, dtype=np.int64).ravel()
    return arr, arr.ctypes.data_as(_lib.i64p), int(arr.size)


def _u64(a):
    arr = np.ascontiguousarray(a, dtype=np.uint64).ravel()
    return arr, arr.ctypes.data_as(_lib.u64p), int(arr.size)


def _f64(a):
    arr = np.ascontiguousarray(a, dtype=np.float64).ravel()
    return arr, arr.ctypes.data_as(_lib.dp), int(arr.size)


def _c128(a):
    arr = np.ascontiguousarray(a, dtype=np.complex128).ravel()
    return arr, arr.ctypes.data_as(_lib.dp), int(arr.size)


class InitializationSettings:
    """Inert mirror of Kokkos::InitializationSettings (Bindings.cpp:854-967); only ``device_id``
    is honoured (it selects the CUDA device of the state vector)."""

    _FIELDS = {"num_threads": 0, "device_id": 0, "map_device_id_by": "", "disable_warnings": False,
               "print_configuration": False, "tune_internals": False, "tools_libs": "",
               "tools_help": False, "tools_args": ""}

    def __init__(self):
        # the reference's binding constructs the settings by calling every setter with its default
        # (Bindings.cpp:855-866), so every has_*() is true from the start
        self._v = dict(self._FIELDS)

    def __getattr__(self, item):
        for prefix in ("get_", "set_", "has_"):
            if item.startswith(prefix) and item[len(prefix):] in self._FIELDS:
                key = item[len(prefix):]
                if prefix == "get_":
                    return lambda: self._v.get(key, self._FIELDS[key])
                if prefix == "has_":
                    return lambda: key in self._v
                def setter(value, _k=key):
                    self._v[_k] = type(self._FIELDS[_k])(value)
                    return self
                return setter
        raise AttributeError(item)

    def __repr__(self):
        lines = ["InitializationSettings:"]
        for k, d in self._FIELDS.items():
            lines.append(f"{k} = {self._v.get(k, d)}")
        return "\n".join(lines)


def kokkos_start():
    """No global runtime to start (the CUDA context is created lazily)."""


def kokkos_end():
    """No global runtime to finalise."""


def kokkos_config_info() -> dict:
    return {"Backend": {"b2sv": _lib.backend_info()}, "Version": lib.b2sv_version().decode()}


def print_configuration():
    print(_lib.backend_info())


class _ObservableBase:
    _dtype_flag = 1

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self._h:
                lib.b2sv_obs_destroy(self._h)
        except Exception:
            pass

    def __repr__(self):
        buf = C.create_string_buffer(1 << 20)
        check(lib.b2sv_obs_name(self._h, buf, len(buf)))
        return buf.value.decode()

    def get_wires(self):
        n = C.c_int(0)
        arr = np.zeros(64, dtype=np.int64)
        check(lib.b2sv_obs_wires(self._h, arr.ctypes.data_as(_lib.i64p), 64, C.byref(n)))
        return [int(x) for x in arr[: n.value]]

    def apply_in_place(self, sv):
        """ObservableKokkos::applyInPlace (ObservablesKokkos.hpp:41): sv <- O sv."""
        check(lib.b2sv_obs_apply(self._h, sv._h))

    def __eq__(self, other):  # Bindings.cpp:608-619: same type and same description
        return type(self) is type(other) and self._key() == other._key()

    def __hash__(self):
        return hash((type(self).__name__, self._key()))


def _make_classes(bits: str, dtype_flag: int, cdtype, rdtype):
    ns = {}

    class NamedObs(_ObservableBase):
        def __init__(self, name, wires):
            self._name, self._wires = str(name), [int(w) for w in wires]
            _, wp, nw = _i64(self._wires)
            h = C.c_void_p()
            check(lib.b2sv_obs_named(self._name.encode(), wp, nw, C.byref(h)))
            super().__init__(h.value)

        def _key(self):
            return (self._name, tuple(self._wires))

    class HermitianObs(_ObservableBase):
        def __init__(self, matrix, wires):
            self._m = np.array(matrix, dtype=np.complex128).ravel()
            self._wires = [int(w) for w in wires]
            _, mp, _n = _c128(self._m)
            _, wp, nw = _i64(self._wires)
            if self._m.size != 4 ** nw:
                raise PLException("Hermitian matrix size does not match the number of wires")
            h = C.c_void_p()
            check(lib.b2sv_obs_hermitian(mp, wp, nw, C.byref(h)))
            super().__init__(h.value)

        def _key(self):
            return (self._m.tobytes(), tuple(self._wires))

    class TensorProdObs(_ObservableBase):
        def __init__(self, obs):
            self._obs = list(obs)
            arr = (C.c_void_p * len(self._obs))(*[o._h for o in self._obs])
            h = C.c_void_p()
            check(lib.b2sv_obs_tensor(arr, len(self._obs), C.byref(h)))
            super().__init__(h.value)

        def _key(self):
            return tuple((type(o).__name__, o._key()) for o in self._obs)

    class Hamiltonian(_ObservableBase):
        def __init__(self, coeffs, obs):
            self._coeffs = np.array(coeffs, dtype=np.float64).ravel()
            self._obs = list(obs)
            if self._coeffs.size != len(self._obs):
                raise PLException("Assertion failed: coeffs_.size() == obs_.size()")
            _, cp, _n = _f64(self._coeffs)
            arr = (C.c_void_p * len(self._obs))(*[o._h for o in self._obs])
            h = C.c_void_p()
            check(lib.b2sv_obs_hamiltonian(cp, arr, len(self._obs), C.byref(h)))
            super().__init__(h.value)

        def _key(self):
            return (self._coeffs.tobytes(),
                    tuple((type(o).__name__, o._key()) for o in self._obs))

    class SparseHamiltonian(_ObservableBase):
        def __init__(self, data, indices, indptr, wires):
            self._data = np.array(data, dtype=np.complex128).ravel()
            self._indices = np.array(indices, dtype=np.uint64).ravel()
            self._indptr = np.array(indptr, dtype=np.uint64).ravel()
            self._wires = [int(w) for w in wires]
            if self._data.size != self._indices.size:
                raise PLException("Assertion failed: data_.size() == indices_.size()")
            _, dptr, nnz = _c128(self._data)
            _, iptr, _ = _u64(self._indices)
            _, pptr, np1 = _u64(self._indptr)
            _, wp, nw = _i64(self._wires)
            h = C.c_void_p()
            check(lib.b2sv_obs_sparse(dptr, iptr, pptr, nnz, np1 - 1, wp, nw, C.byref(h)))
            super().__init__(h.value)

        def _key(self):
            return (self._data.tobytes(), self._indices.tobytes(), self._indptr.tobytes())

    class OpsStruct:
        """OpsData<P> (reference AdjointDiffKokkos.hpp:17-173)."""

        def __init__(self, names, params, wires, inverses, matrices=None):
            n = len(names)
            if not (len(params) == n and len(wires) == n and len(inverses) == n):
                raise PLException("Incompatible number of ops, params, wires and inverses")
            # marshalling is on the e2e path of every apply(): flat numpy buffers, no per-gate arrays
            self.names = [s if type(s) is str else str(s) for s in names]
            self.params = [p if type(p) is list else np.asarray(p, dtype=np.float64).ravel().tolist()
                           for p in params]
            self.wires = [ws if type(ws) is list else [int(w) for w in ws] for ws in wires]
            self.inverses = [bool(i) for i in inverses]
            self._mats = None
            mat_ptrs = None
            if matrices is not None and any(m is not None and np.size(m) for m in matrices):
                mats = list(matrices) + [None] * (n - len(matrices))
                self._mats = [None if m is None or np.size(m) == 0
                              else np.ascontiguousarray(m, dtype=np.complex128).ravel() for m in mats]
                mat_ptrs = (_lib.dp * n)(*[
                    m.ctypes.data_as(_lib.dp) if m is not None else C.cast(None, _lib.dp)
                    for m in self._mats])
            c_names = (C.c_char_p * n)(*[s.encode() for s in self.names])
            flat_p = np.fromiter(_chain(self.params), dtype=np.float64)
            nparams = np.fromiter(map(len, self.params), dtype=np.int32, count=n)
            flat_w = np.fromiter(_chain(self.wires), dtype=np.int64)
            nwires = np.fromiter(map(len, self.wires), dtype=np.int32, count=n)
            inv = np.fromiter(self.inverses, dtype=np.int32, count=n)
            h = C.c_void_p()
            check(lib.b2sv_ops_create(n, c_names, flat_p.ctypes.data_as(_lib.dp),
                                      nparams.ctypes.data_as(_lib.ip), flat_w.ctypes.data_as(_lib.i64p),
                                      nwires.ctypes.data_as(_lib.ip), inv.ctypes.data_as(_lib.ip),
                                      mat_ptrs, C.byref(h)))
            self._h = h

        def __del__(self):
            try:
                if self._h:
                    lib.b2sv_ops_destroy(self._h)
            except Exception:
                pass

        def plan(self, num_qubits):
            """Host-only: what the fusion scheduler does with this op list on ``num_qubits`` wires
            (b2sv_plan_ops).  Works without a GPU."""
            v = [C.c_uint64() for _ in range(5)]
            check(lib.b2sv_plan_ops(self._h, int(num_qubits), dtype_flag, *[C.byref(x) for x in v]))
            return dict(zip(("passes", "rounds", "arithmetic_ops", "absorbed_perms", "fused_stores"),
                            (x.value for x in v)))

        def plan_sharded(self, num_qubits, world, with_text=False):
            """Host-only: runs / exchanges / tile passes of this op list on a state sharded over
            ``world`` ranks (b2sv_plan_sharded). Works without a GPU."""
            v = (C.c_uint64 * 5)()
            buf = C.create_string_buffer(64 << 20) if with_text else None
            check(lib.b2sv_plan_sharded(self._h, int(num_qubits), int(world), dtype_flag, v, buf,
                                        len(buf) if buf else 0))
            out = dict(zip(("runs", "exchanges", "passes", "exchanged_bits", "bytes_per_rank"),
                           (int(x) for x in v)))
            if with_text:
                out["text"] = buf.value.decode()
            return out

        def __len__(self):
            return len(self.names)

        def __repr__(self):  # Bindings.cpp:749-762
            parts = []
            for nm, p, inv in zip(self.names, self.params, self.inverses):
                parts.append("{'name': %s, 'params': %s, 'inv': %d}" % (nm, list(p), int(inv)))
            return "Operations: [" + ",".join(parts) + "]"

    class LightningKokkos:
        """StateVectorKokkos<P> + MeasuresKokkos<P> as bound by Bindings.cpp:59-585."""

        def __init__(self, arg, settings: InitializationSettings | None = None):
            device_id = settings.get_device_id() if settings is not None else 0
            self._h = C.c_void_p()
            if isinstance(arg, (int, np.integer)):
                check(lib.b2sv_create(int(arg), dtype_flag, device_id, C.byref(self._h)))
            else:  # Bindings.cpp:71-85: from a host array of 2^n amplitudes
                arr = np.ascontiguousarray(arg, dtype=cdtype).ravel()
                n = int(arr.size).bit_length() - 1
                if arr.size == 0 or (1 << n) != arr.size:
                    raise PLException("state vector length must be a power of two")
                check(lib.b2sv_create(n, dtype_flag, device_id, C.byref(self._h)))
                self.HostToDevice(arr)

        @classmethod
        def sharded(cls, num_qubits, device_id, rank, world, nccl_unique_id: bytes):
            """State of ``num_qubits`` wires sharded over ``world`` GPUs (one process per GPU):
            rank = top log2(world) index bits = wires 0..g-1.  ``nccl_unique_id`` is the 128-byte
            id from :func:`pennylane_lightning_kokkos_b200.dist.nccl_unique_id`, identical on all ranks.
            (New: the reference is single-device)"""
            self = cls.__new__(cls)
            self._h = C.c_void_p()
            buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            check(lib.b2sv_create_sharded(int(num_qubits), dtype_flag, int(device_id), int(rank),
                                          int(world), C.cast(buf, C.c_void_p), C.byref(self._h)))
            return self

        def __del__(self):
            try:
                if self._h:
                    lib.b2sv_destroy(self._h)
                    self._h = None
            except Exception:
                pass

        # -- state management
        def setBasisState(self, index):
            check(lib.b2sv_set_basis_state(self._h, int(index)))

        def setStateVector(self, indices, state):
            _, ip_, n = _u64(indices)
            vals, vp_, nv = _c128(state)
            if n != nv:
                raise PLException("indices and state must have the same length")
            check(lib.b2sv_set_state_vector(self._h, ip_, vp_, n))

        def setStateOnWires(self, wires, state):
            """State preparation on a subset of wires, index table built on the device
            (replaces lightning_kokkos.py:317-327 + setStateVector)."""
            _, wp, nw = _i64(wires)
            vals, vp_, nv = _c128(state)
            if nv != 1 << nw:
                raise PLException("state must have 2**len(wires) amplitudes")
            check(lib.b2sv_set_state_on_wires(self._h, wp, nw, vp_))

        def DeviceToHost(self, host_sv):
            if not (isinstance(host_sv, np.ndarray) and host_sv.dtype == cdtype
                    and host_sv.flags["C_CONTIGUOUS"]):
                # the reference silently copies a wrong-dtype array (py::array::forcecast) and the
                # result is lost (SURVEY 8b "Ownership"); we refuse instead
                raise PLException(f"DeviceToHost needs a C-contiguous {np.dtype(cdtype)} array")
            if host_sv.size:
                check(lib.b2sv_d2h(self._h, host_sv.ctypes.data_as(C.c_void_p), host_sv.size))

        def HostToDevice(self, host_sv, length=None):
            arr = np.ascontiguousarray(host_sv, dtype=cdtype).ravel()
            n = int(length) if length is not None else arr.size
            if n:
                check(lib.b2sv_h2d(self._h, arr.ctypes.data_as(C.c_void_p), n))

        def numQubits(self):
            n = C.c_int()
            check(lib.b2sv_num_qubits(self._h, C.byref(n)))
            return n.value

        def dataLength(self):
            n = C.c_uint64()
            check(lib.b2sv_data_length(self._h, C.byref(n)))
            return n.value

        def resetKokkos(self):
            check(lib.b2sv_reset(self._h))

        # -- gates
        def _named(self, name, wires, adjoint, params):
            _, wp, nw = _i64(wires)
            _, pp, npar = _f64(params if params is not None else [])
            check(lib.b2sv_apply(self._h, name.encode(), wp, nw, int(bool(adjoint)), pp, npar))

        def apply(self, *args):
            """The three overloads of Bindings.cpp:233-261."""
            if len(args) in (3, 4) and not isinstance(args[0], str):
                names, wires, adjoints = args[0], args[1], args[2]
                params = args[3] if len(args) == 4 else [[] for _ in names]
                if len(names) != len(wires):
                    raise PLException("Incompatible number of ops and wires")
                if len(names) != len(adjoints):
                    raise PLException("Incompatible number of ops and adjoints")
                ops = OpsStruct(names, params, wires, adjoints)
                check(lib.b2sv_apply_ops(self._h, ops._h, 0))
                return
            if len(args) == 5 and isinstance(args[0], str):
                name, wires, inv, _params, matrix = args
                m = np.asarray(matrix)
                if m.size == 0 or name in _GATES_1:
                    return self._named(name, wires, inv, [])
                _, wp, nw = _i64(wires)
                _, mp, nm = _c128(m)
                if nm != 4 ** nw:
                    raise PLException("matrix size does not match the number of wires")
                check(lib.b2sv_apply_matrix(self._h, wp, nw, int(bool(inv)), mp))
                return
            raise TypeError("apply(): incompatible function arguments")

        def apply_ops(self, ops: "OpsStruct", adjoint=False):
            """Fast path: the whole list in one call so the scheduler can fuse."""
            check(lib.b2sv_apply_ops(self._h, ops._h, int(bool(adjoint))))

        def applyGenerator(self, name, wires, adjoint=False, params=None):
            _, wp, nw = _i64(wires)
            s = C.c_double()
            check(lib.b2sv_apply_generator(self._h, name.encode(), wp, nw, int(bool(adjoint)),
                                           C.byref(s)))
            return s.value

        # -- measurements
        def ExpectationValue(self, *args):
            """The four overloads of Bindings.cpp:434-516."""
            out = C.c_double()
            if len(args) == 4 and isinstance(args[0], str):
                name, wires, _params, matrix = args
                m = np.asarray(matrix)
                if name in ("Identity", "PauliX", "PauliY", "PauliZ", "Hadamard"):
                    # MK.hpp:84-95: wires are reversed when a matrix is supplied (1-wire: no-op)
                    w = list(wires)[::-1] if m.size else list(wires)
                    _, wp, nw = _i64(w)
                    check(lib.b2sv_expval_named(self._h, name.encode(), wp, nw, C.byref(out)))
                    return out.value
                return self.ExpectationValue(wires, m)
            if len(args) == 4:  # list of names: "#"+concat is never a named observable
                _names, wires, _params, matrix = args
                return self.ExpectationValue(wires, np.asarray(matrix))
            if len(args) == 2:
                wires, matrix = args
                _, wp, nw = _i64(wires)
                _, mp, nm = _c128(matrix)
                if nm != 4 ** nw:
                    raise PLException("matrix size does not match the number of wires")
                check(lib.b2sv_expval_matrix(self._h, wp, nw, mp, C.byref(out)))
                return out.value
            if len(args) == 3:
                data, indices, indptr = args
                _, dptr, nnz = _c128(data)
                _, iptr, _n = _u64(indices)
                _, pptr, np1 = _u64(indptr)
                check(lib.b2sv_expval_csr(self._h, dptr, iptr, pptr, nnz, np1 - 1, C.byref(out)))
                return out.value
            raise TypeError("ExpectationValue(): incompatible function arguments")

        def expval_z_all(self):
            """<Z_w> for every wire w from one read pass over the state."""
            n = self.numQubits()
            out = np.zeros(n, dtype=np.float64)
            check(lib.b2sv_expval_z_all(self._h, out.ctypes.data_as(_lib.dp), n))
            return out

        def expval(self, obs) -> float:
            """MeasuresKokkos::expval(Observable) (MK.hpp:354-360; unbound in the reference)."""
            out = C.c_double()
            check(lib.b2sv_expval_obs(self._h, obs._h, C.byref(out)))
            return out.value

        def var(self, obs) -> float:
            """MeasuresKokkos::var(Observable) (MK.hpp:368-381; unbound in the reference)."""
            out = C.c_double()
            check(lib.b2sv_var_obs(self._h, obs._h, C.byref(out)))
            return out.value

        def probs(self, wires):
            wires = [int(w) for w in wires]
            nq = self.numQubits()
            m = len(wires) if wires else nq
            out = np.zeros(1 << m, dtype=np.float64)
            _, wp, nw = _i64(wires)
            check(lib.b2sv_probs(self._h, wp, nw, out.ctypes.data_as(_lib.dp)))
            return out.astype(rdtype, copy=False)

        def GenerateSamples(self, num_wires, num_shots, seed=5374857):
            nq = self.numQubits()
            out = np.zeros((int(num_shots), nq), dtype=np.uint64)
            check(lib.b2sv_generate_samples(self._h, int(num_shots), int(seed),
                                            out.ctypes.data_as(_lib.u64p)))
            return out.reshape(int(num_shots), int(num_wires))

        # -- linear algebra (util/LinearAlgebraKokkos.hpp:30-61,155-236) and copies (SV.hpp:550-554,1596)
        def clone(self):
            new = type(self).__new__(type(self))
            new._h = C.c_void_p()
            check(lib.b2sv_clone(self._h, C.byref(new._h)))
            return new

        def updateData(self, other):
            check(lib.b2sv_copy(self._h, other._h))

        def inner_product(self, other) -> complex:
            """<self|other>"""
            re, im = C.c_double(), C.c_double()
            check(lib.b2sv_inner_product(self._h, other._h, C.byref(re), C.byref(im)))
            return complex(re.value, im.value)

        def axpy(self, alpha, x):
            """self += alpha * x"""
            a = complex(alpha)
            check(lib.b2sv_axpy(a.real, a.imag, x._h, self._h))

        # -- extras used by tests / bench
        def set_fusion(self, fuse: bool):
            check(lib.b2sv_set_fusion(self._h, int(bool(fuse))))

        def stats(self):
            s, l = C.c_uint64(), C.c_uint64()
            check(lib.b2sv_get_stats(self._h, C.byref(s), C.byref(l)))
            return {"sweeps": s.value, "launches": l.value}

        def last_adjoint_traffic(self):
            b = C.c_uint64()
            check(lib.b2sv_last_adjoint_traffic(self._h, C.byref(b)))
            return b.value

        def reset_stats(self):
            check(lib.b2sv_reset_stats(self._h))

        def comm_stats(self):
            sw, by, peer = C.c_uint64(), C.c_uint64(), C.c_int()
            check(lib.b2sv_comm_stats(self._h, C.byref(sw), C.byref(by), C.byref(peer)))
            return {"swaps": sw.value, "swap_bytes_per_rank": by.value,
                    "path": "nvlink-peer-kernel" if peer.value else "nccl-sendrecv"}

        def last_upload_bytes(self):
            b = C.c_uint64()
            check(lib.b2sv_last_upload_bytes(self._h, C.byref(b)))
            return b.value

        def amplitudes(self, indices):
            """Sampled read: complex128 amplitudes at the given global flat indices."""
            _, ip_, n = _u64(indices)
            out = np.zeros(n, dtype=np.complex128)
            check(lib.b2sv_get_amplitudes(self._h, ip_, n, out.ctypes.data_as(_lib.dp)))
            return out

        def trace_begin(self):
            check(lib.b2sv_trace_begin(self._h))

        def trace_end(self, cap=65536):
            """[(kind, start_ms, dur_ms)]: kind 0 tile pass, 1 matrix kernel, 2 exchange."""
            kinds = np.zeros(cap, dtype=np.int32)
            t0 = np.zeros(cap, dtype=np.float64)
            dt = np.zeros(cap, dtype=np.float64)
            n = C.c_int(0)
            check(lib.b2sv_trace_end(self._h, kinds.ctypes.data_as(_lib.ip),
                                     t0.ctypes.data_as(_lib.dp), dt.ctypes.data_as(_lib.dp), cap,
                                     C.byref(n)))
            m = min(n.value, cap)
            return [(int(kinds[i]), float(t0[i]), float(dt[i])) for i in range(m)]

        def layout(self):
            """Logical index bit q currently lives at physical bit layout()[q] (sharded states move
            qubits between rank bits and shard-local bits lazily)."""
            arr = np.zeros(64, dtype=np.int32)
            n = C.c_int(0)
            check(lib.b2sv_layout(self._h, arr.ctypes.data_as(_lib.ip), 64, C.byref(n)))
            return [int(x) for x in arr[: n.value]]

        def normalize_layout(self):
            check(lib.b2sv_normalize_layout(self._h))

        def sync(self):
            check(lib.b2sv_sync(self._h))

        def device_ptr(self) -> int:
            p = C.c_void_p()
            check(lib.b2sv_device_ptr(self._h, C.byref(p)))
            return p.value

        def stream_ptr(self) -> int:
            p = C.c_void_p()
            check(lib.b2sv_stream(self._h, C.byref(p)))
            return p.value or 0

    def _gate_method(name):
        def method(self, wires, adjoint=False, params=None):
            return self._named(name, wires, adjoint, params)
        method.__name__ = name
        method.__doc__ = f"Apply the {name} gate (Bindings.cpp:109-433)."
        return method

    for g in _GATES_1:
        setattr(LightningKokkos, g, _gate_method(g))

    class AdjointJacobian:
        """AdjointJacobianKokkos<P> as bound by Bindings.cpp:768-821."""

        def create_ops_list(self, names, params, wires, inverses, matrices):
            return OpsStruct(names, params, wires, inverses, matrices)

        def adjoint_jacobian(self, sv, observables, operations, trainable_params):
            tp = np.ascontiguousarray(trainable_params, dtype=np.uint64).ravel()
            n_obs = len(observables)
            jac = np.zeros((n_obs, tp.size), dtype=np.float64)
            oarr = (C.c_void_p * n_obs)(*[o._h for o in observables])
            check(lib.b2sv_adjoint_jacobian(sv._h, oarr, n_obs, operations._h,
                                            tp.ctypes.data_as(_lib.u64p), int(tp.size),
                                            jac.ctypes.data_as(_lib.dp)))
            return jac.astype(rdtype, copy=False)

        def vjp(self, sv, observables, operations, trainable_params, dy):
            """Vector-Jacobian product ``dy @ jac`` the way the reference device computes it
            (lightning_kokkos.py:689-727): the adjoint Jacobian of the one Hamiltonian
            ``sum_i dy[i] * observables[i]`` -- a single reverse sweep."""
            dy = np.asarray(dy)
            if np.iscomplexobj(dy):  # lightning_kokkos.py:709-712
                raise ValueError("The vjp method only works with a real-valued dy when the tape is "
                                 "returning an expectation value")
            if dy.size != len(observables):  # lightning_kokkos.py:704-707
                raise ValueError("Number of observables in the tape must be the same as the length "
                                 "of dy in the vjp method")
            tp = np.ascontiguousarray(trainable_params, dtype=np.uint64).ravel()
            n_obs = len(observables)
            out = np.zeros(tp.size, dtype=np.float64)
            oarr = (C.c_void_p * n_obs)(*[o._h for o in observables])
            _, dyp, _n = _f64(dy.astype(np.float64).ravel())
            check(lib.b2sv_adjoint_vjp(sv._h, oarr, n_obs, dyp, operations._h,
                                       tp.ctypes.data_as(_lib.u64p), int(tp.size),
                                       out.ctypes.data_as(_lib.dp)))
            return out.astype(rdtype, copy=False)

    for cls, nm in ((NamedObs, "NamedObsKokkos"), (HermitianObs, "HermitianObsKokkos"),
                    (TensorProdObs, "TensorProdObsKokkos"), (Hamiltonian, "HamiltonianKokkos"),
                    (SparseHamiltonian, "SparseHamiltonianKokkos"), (OpsStruct, "OpsStructKokkos"),
                    (LightningKokkos, "LightningKokkos"),
                    (AdjointJacobian, "AdjointJacobianKokkos")):
        cls.__name__ = cls.__qualname__ = f"{nm}_C{bits}"
        ns[cls.__name__] = cls
    return ns


globals().update(_make_classes("64", 0, np.complex64, np.float32))
globals().update(_make_classes("128", 1, np.complex128, np.float64))
ObservableKokkos_C64 = _ObservableBase
ObservableKokkos_C128 = _ObservableBase
