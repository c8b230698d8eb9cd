"""ctypes loader for libb2sv.so -- the C-ABI engine (include/b2sv.h).

There is no fallback of any kind: if the shared library is missing the import fails, and if
no CUDA device is present every compute entry point raises PLException.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2SV_LIB: developer override (kernel A/B experiments load a differently built libb2sv)
LIB_PATH = os.environ.get("B2SV_LIB") or os.path.join(_HERE, "libb2sv.so")


class PLException(RuntimeError):
    """Mirror of the reference's PLException (reference Bindings.cpp:837, util/Error.hpp)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C pennylane_lightning_kokkos_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). b2sv has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

vp = C.c_void_p
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

# every symbol include/b2sv.h declares, with its argument types
PROTOTYPES = {
    "b2sv_last_error": (C.c_char_p, []),
    "b2sv_version": (C.c_char_p, []),
    "b2sv_backend_info": (C.c_int, [C.c_char_p, C.c_size_t]),
    "b2sv_device_count": (C.c_int, [ip]),
    "b2sv_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "b2sv_create_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]),
    "b2sv_comm_unique_id": (C.c_int, [vp]),
    "b2sv_destroy": (C.c_int, [vp]),
    "b2sv_clone": (C.c_int, [vp, C.POINTER(vp)]),
    "b2sv_copy": (C.c_int, [vp, vp]),
    "b2sv_reset": (C.c_int, [vp]),
    "b2sv_init_zeros": (C.c_int, [vp]),
    "b2sv_set_basis_state": (C.c_int, [vp, C.c_uint64]),
    "b2sv_set_state_vector": (C.c_int, [vp, u64p, dp, C.c_size_t]),
    "b2sv_set_state_on_wires": (C.c_int, [vp, i64p, C.c_int, dp]),
    "b2sv_h2d": (C.c_int, [vp, vp, C.c_size_t]),
    "b2sv_d2h": (C.c_int, [vp, vp, C.c_size_t]),
    "b2sv_get_amplitudes": (C.c_int, [vp, u64p, C.c_size_t, dp]),
    "b2sv_trace_begin": (C.c_int, [vp]),
    "b2sv_trace_end": (C.c_int, [vp, ip, dp, dp, C.c_int, ip]),
    "b2sv_num_qubits": (C.c_int, [vp, ip]),
    "b2sv_data_length": (C.c_int, [vp, u64p]),
    "b2sv_device_ptr": (C.c_int, [vp, C.POINTER(vp)]),
    "b2sv_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "b2sv_sync": (C.c_int, [vp]),
    "b2sv_apply": (C.c_int, [vp, C.c_char_p, i64p, C.c_int, C.c_int, dp, C.c_int]),
    "b2sv_apply_matrix": (C.c_int, [vp, i64p, C.c_int, C.c_int, dp]),
    "b2sv_apply_ops": (C.c_int, [vp, vp, C.c_int]),
    "b2sv_apply_generator": (C.c_int, [vp, C.c_char_p, i64p, C.c_int, C.c_int, dp]),
    "b2sv_set_fusion": (C.c_int, [vp, C.c_int]),
    "b2sv_get_stats": (C.c_int, [vp, u64p, u64p]),
    "b2sv_plan_ops": (C.c_int, [vp, C.c_int, C.c_int, u64p, u64p, u64p, u64p, u64p]),
    "b2sv_plan_sharded": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, u64p, C.c_char_p, C.c_size_t]),
    "b2sv_comm_stats": (C.c_int, [vp, u64p, u64p, ip]),
    "b2sv_last_upload_bytes": (C.c_int, [vp, u64p]),
    "b2sv_normalize_layout": (C.c_int, [vp]),
    "b2sv_layout": (C.c_int, [vp, ip, C.c_int, ip]),
    "b2sv_reset_stats": (C.c_int, [vp]),
    "b2sv_last_adjoint_traffic": (C.c_int, [vp, u64p]),
    "b2sv_debug_tile_prof": (C.c_int, [u64p]),
    "b2sv_ops_create": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), dp, ip, i64p, ip, ip,
                                  C.POINTER(dp), C.POINTER(vp)]),
    "b2sv_ops_destroy": (C.c_int, [vp]),
    "b2sv_ops_size": (C.c_int, [vp, ip, ip]),
    "b2sv_expval_named": (C.c_int, [vp, C.c_char_p, i64p, C.c_int, dp]),
    "b2sv_expval_z_all": (C.c_int, [vp, dp, C.c_int]),
    "b2sv_invalidate": (C.c_int, [vp]),
    "b2sv_expval_matrix": (C.c_int, [vp, i64p, C.c_int, dp, dp]),
    "b2sv_expval_csr": (C.c_int, [vp, dp, u64p, u64p, C.c_size_t, C.c_size_t, dp]),
    "b2sv_csr_create": (C.c_int, [vp, dp, u64p, u64p, C.c_size_t, C.c_size_t, C.POINTER(vp)]),
    "b2sv_csr_destroy": (C.c_int, [vp]),
    "b2sv_expval_csr_resident": (C.c_int, [vp, vp, dp]),
    "b2sv_expval_obs": (C.c_int, [vp, vp, dp]),
    "b2sv_var_obs": (C.c_int, [vp, vp, dp]),
    "b2sv_probs": (C.c_int, [vp, i64p, C.c_int, dp]),
    "b2sv_generate_samples": (C.c_int, [vp, C.c_size_t, C.c_uint64, u64p]),
    "b2sv_inner_product": (C.c_int, [vp, vp, dp, dp]),
    "b2sv_axpy": (C.c_int, [C.c_double, C.c_double, vp, vp]),
    "b2sv_obs_named": (C.c_int, [C.c_char_p, i64p, C.c_int, C.POINTER(vp)]),
    "b2sv_obs_hermitian": (C.c_int, [dp, i64p, C.c_int, C.POINTER(vp)]),
    "b2sv_obs_tensor": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(vp)]),
    "b2sv_obs_hamiltonian": (C.c_int, [dp, C.POINTER(vp), C.c_int, C.POINTER(vp)]),
    "b2sv_obs_sparse": (C.c_int, [dp, u64p, u64p, C.c_size_t, C.c_size_t, i64p, C.c_int,
                                  C.POINTER(vp)]),
    "b2sv_obs_destroy": (C.c_int, [vp]),
    "b2sv_obs_name": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
    "b2sv_obs_wires": (C.c_int, [vp, i64p, C.c_int, ip]),
    "b2sv_obs_apply": (C.c_int, [vp, vp]),
    "b2sv_adjoint_jacobian": (C.c_int, [vp, C.POINTER(vp), C.c_int, vp, u64p, C.c_int, dp]),
    "b2sv_adjoint_vjp": (C.c_int, [vp, C.POINTER(vp), C.c_int, dp, vp, u64p, C.c_int, dp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)  # AttributeError here = the library does not export the ABI
    _f.restype = _res
    _f.argtypes = _args


def check(rc: int):
    if rc != 0:
        raise PLException(lib.b2sv_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib.b2sv_device_count(C.byref(n)))
    return n.value


def backend_info() -> str:
    buf = C.create_string_buffer(4096)
    check(lib.b2sv_backend_info(buf, len(buf)))
    return buf.value.decode()
