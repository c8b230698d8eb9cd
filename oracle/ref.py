"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/_ref/libref_oracle.so.

libref_oracle.so is the UNMODIFIED reference (PennyLane-Lightning-Kokkos headers, reached
by include path from /root/reference) compiled against oracle/kokkos_shim with OpenMP;
see oracle/Makefile and oracle/ref_capi.cpp.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module: it is the checker
and the timed CPU baseline, never part of the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
_lib = None

c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)


def available() -> bool:
    return os.path.exists(_SO)


def build(force: bool = False) -> bool:
    """Compile the reference oracle if /root/reference is present. Returns availability."""
    ref = "/root/reference/pennylane_lightning_kokkos/src"
    if os.path.isdir(ref) and (force or not os.path.exists(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return available()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{_SO} missing: run `make -C oracle` in the build container "
                "(needs /root/reference)")
        L = C.CDLL(_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_sv_create.restype = C.c_void_p
        for f in ("ref_obs_named", "ref_obs_hermitian", "ref_obs_tensor",
                  "ref_obs_hamiltonian", "ref_obs_sparse"):
            getattr(L, f).restype = C.c_void_p
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())


def _wires(w):
    a = np.ascontiguousarray(w, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p), int(a.size)


def _dbl(x):
    a = np.ascontiguousarray(x, dtype=np.float64).ravel()
    return a, a.ctypes.data_as(c_dp)


def _cplx(x):
    a = np.ascontiguousarray(x, dtype=np.complex128).ravel()
    return a, a.ctypes.data_as(c_dp)


def num_threads() -> int:
    return lib().ref_num_threads()


def set_num_threads(n: int):
    lib().ref_set_num_threads(int(n))


class RefObs:
    """Reference observable object (reference ObservablesKokkos.hpp)."""

    def __init__(self, handle, prec, keep=()):
        if not handle:
            raise RuntimeError(lib().ref_last_error().decode())
        self.h = C.c_void_p(handle)
        self.prec = prec
        self._keep = keep

    def __del__(self):
        try:
            lib().ref_obs_destroy(self.h)
        except Exception:
            pass

    def name(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        _chk(lib().ref_obs_name(self.h, buf, len(buf)))
        return buf.value.decode()

    @staticmethod
    def named(name, wires, prec=1):
        _, wp, nw = _wires(wires)
        return RefObs(lib().ref_obs_named(prec, name.encode(), wp, nw), prec)

    @staticmethod
    def hermitian(matrix, wires, prec=1):
        _, mp = _cplx(matrix)
        _, wp, nw = _wires(wires)
        return RefObs(lib().ref_obs_hermitian(prec, mp, wp, nw), prec)

    @staticmethod
    def tensor(obs, prec=1):
        arr = (C.c_void_p * len(obs))(*[o.h for o in obs])
        return RefObs(lib().ref_obs_tensor(prec, arr, len(obs)), prec, tuple(obs))

    @staticmethod
    def hamiltonian(coeffs, obs, prec=1):
        _, cp = _dbl(coeffs)
        arr = (C.c_void_p * len(obs))(*[o.h for o in obs])
        return RefObs(lib().ref_obs_hamiltonian(prec, cp, arr, len(obs)), prec, tuple(obs))

    @staticmethod
    def sparse(data, indices, indptr, wires, prec=1):
        _, dp = _cplx(data)
        ia, ip, nnz = _wires(indices)
        pa, pp, np1 = _wires(indptr)
        _, wp, nw = _wires(wires)
        return RefObs(lib().ref_obs_sparse(prec, dp, ip, pp, nnz, np1 - 1, wp, nw), prec)


class RefStateVector:
    """Reference StateVectorKokkos<P> + MeasuresKokkos<P> (P = float for complex64)."""

    def __init__(self, num_qubits: int, dtype=np.complex128):
        self.dtype = np.dtype(dtype)
        self.prec = 1 if self.dtype == np.complex128 else 0
        self.n = num_qubits
        h = lib().ref_sv_create(self.prec, num_qubits)
        if not h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            lib().ref_sv_destroy(self.h)
        except Exception:
            pass

    # -- state management (reference StateVectorKokkos.hpp:476-532,1596-1636)
    def reset(self):
        _chk(lib().ref_sv_reset(self.h))

    def set_basis_state(self, index):
        _chk(lib().ref_sv_set_basis_state(self.h, C.c_int64(index)))

    def set_state_vector(self, indices, values):
        ia, ip, n = _wires(indices)
        _, vp = _cplx(values)
        _chk(lib().ref_sv_set_state_vector(self.h, ip, vp, C.c_int64(n)))

    def h2d(self, state):
        a = np.ascontiguousarray(state, dtype=self.dtype).ravel()
        assert a.size == 1 << self.n
        _chk(lib().ref_sv_h2d(self.h, a.ctypes.data_as(C.c_void_p), C.c_int64(a.size)))

    def d2h(self):
        out = np.empty(1 << self.n, dtype=self.dtype)
        _chk(lib().ref_sv_d2h(self.h, out.ctypes.data_as(C.c_void_p), C.c_int64(out.size)))
        return out

    def amplitudes(self, indices):
        """Amplitudes at the given flat indices (sampled read, no full copy)."""
        ia, ip, n = _wires(indices)
        out = np.empty(n, dtype=np.complex128)
        _chk(lib().ref_sv_get_amplitudes(self.h, ip, C.c_int64(n), out.ctypes.data_as(c_dp)))
        return out

    # -- gates
    def apply(self, name, wires, inverse=False, params=()):
        _, wp, nw = _wires(wires)
        pa, pp = _dbl(params)
        _chk(lib().ref_sv_apply(self.h, name.encode(), wp, nw, int(bool(inverse)), pp, pa.size))

    def apply_matrix(self, matrix, wires, inverse=False):
        _, wp, nw = _wires(wires)
        _, mp = _cplx(matrix)
        _chk(lib().ref_sv_apply_matrix(self.h, wp, nw, int(bool(inverse)), mp))

    def apply_ops(self, ops):
        for name, wires, inverse, params in ops:
            self.apply(name, wires, inverse, params)

    def apply_generator(self, name, wires, adj=False) -> float:
        _, wp, nw = _wires(wires)
        s = C.c_double()
        _chk(lib().ref_sv_apply_generator(self.h, name.encode(), wp, nw, int(bool(adj)),
                                          C.byref(s)))
        return s.value

    # -- measurements
    def expval_named(self, name, wires) -> float:
        _, wp, nw = _wires(wires)
        out = C.c_double()
        _chk(lib().ref_expval_named(self.h, name.encode(), wp, nw, C.byref(out)))
        return out.value

    def expval_matrix(self, matrix, wires) -> float:
        _, wp, nw = _wires(wires)
        _, mp = _cplx(matrix)
        out = C.c_double()
        _chk(lib().ref_expval_matrix(self.h, wp, nw, mp, C.byref(out)))
        return out.value

    def expval_csr(self, data, indices, indptr) -> float:
        _, dp = _cplx(data)
        ia, ip, nnz = _wires(indices)
        pa, pp, np1 = _wires(indptr)
        out = C.c_double()
        _chk(lib().ref_expval_csr(self.h, dp, ip, pp, C.c_int64(nnz), C.c_int64(np1 - 1),
                                  C.byref(out)))
        return out.value

    def expval_obs(self, obs: RefObs) -> float:
        out = C.c_double()
        _chk(lib().ref_expval_obs(self.h, obs.h, C.byref(out)))
        return out.value

    def var_obs(self, obs: RefObs) -> float:
        out = C.c_double()
        _chk(lib().ref_var_obs(self.h, obs.h, C.byref(out)))
        return out.value

    def apply_obs(self, obs: RefObs):
        _chk(lib().ref_obs_apply(self.h, obs.h))

    def probs(self, wires=None):
        if wires is None:
            out = np.empty(1 << self.n, dtype=np.float64)
            _chk(lib().ref_probs(self.h, None, 0, 1, out.ctypes.data_as(c_dp)))
            return out
        _, wp, nw = _wires(wires)
        out = np.empty(1 << nw, dtype=np.float64)
        _chk(lib().ref_probs(self.h, wp, nw, 0, out.ctypes.data_as(c_dp)))
        return out

    def generate_samples(self, shots):
        out = np.empty((shots, self.n), dtype=np.uint64)
        _chk(lib().ref_generate_samples(self.h, C.c_int64(shots),
                                        out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    # -- adjoint Jacobian (reference AdjointDiffKokkos.hpp:404-478)
    def adjoint_jacobian(self, observables, ops, trainable_params):
        """ops: list of (name, wires, inverse, params); returns jac[n_obs, n_tp]."""
        nops = len(ops)
        names = (C.c_char_p * nops)(*[o[0].encode() for o in ops])
        params = np.array([p for o in ops for p in o[3]], dtype=np.float64)
        nparams = (C.c_int * nops)(*[len(o[3]) for o in ops])
        wires = np.array([w for o in ops for w in o[1]], dtype=np.int64)
        nwires = (C.c_int * nops)(*[len(o[1]) for o in ops])
        inv = (C.c_int * nops)(*[int(bool(o[2])) for o in ops])
        tp = np.ascontiguousarray(trainable_params, dtype=np.int64)
        oarr = (C.c_void_p * len(observables))(*[o.h for o in observables])
        jac = np.zeros((len(observables), tp.size), dtype=np.float64)
        _chk(lib().ref_adjoint_jacobian(
            self.h, oarr, len(observables), nops, names,
            params.ctypes.data_as(c_dp), nparams, wires.ctypes.data_as(c_i64p), nwires, inv,
            tp.ctypes.data_as(c_i64p), int(tp.size), jac.ctypes.data_as(c_dp)))
        return jac
