"""ctypes loader for the pass emulator (csrc/tools/pass_emulator.cpp): TEST INFRASTRUCTURE that
re-executes the scheduler's pass descriptors on the CPU exactly as the tile executor would.
Built on demand into tests/_emu/ (git-ignored)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pennylane_lightning_kokkos_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_emu", "libb2sv_emu.so")
SRCS = [os.path.join(CSRC, "tools", "pass_emulator.cpp"), os.path.join(CSRC, "schedule.cpp"),
        os.path.join(CSRC, "gates.cpp")]


def build():
    deps = SRCS + [os.path.join(CSRC, h) for h in ("schedule.hpp", "ir.hpp", "common.hpp")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                           f"-I{CSRC}", "-o", OUT] + SRCS)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.b2emu_last_error.restype = C.c_char_p
    return _lib


def run(circ, n, psi, f32=False, B=None, R=None, low=5, max_heavy=8, factor=True, store_mode=1):
    """Apply `circ` [(name, wires, inverse, params)] to psi through schedule + emulated passes."""
    B = B or (13 if f32 else 12)
    R = R or (5 if f32 else 4)
    names = (C.c_char_p * len(circ))(*[c[0].encode() for c in circ])
    wires = np.array([w for c in circ for w in c[1]] + [0], dtype=np.int64)
    nw = np.array([len(c[1]) for c in circ] + [0], dtype=np.int32)
    inv = np.array([int(c[2]) for c in circ] + [0], dtype=np.int32)
    params = np.array([p for c in circ for p in c[3]] + [0.0], dtype=np.float64)
    npar = np.array([len(c[3]) for c in circ] + [0], dtype=np.int32)
    st = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    stats = np.zeros(6, dtype=np.uint64)
    rc = lib().b2emu_run(n, int(f32), B, R, low, max_heavy, int(factor), store_mode, len(circ), names,
                         wires.ctypes.data_as(C.c_void_p), nw.ctypes.data_as(C.c_void_p),
                         inv.ctypes.data_as(C.c_void_p), params.ctypes.data_as(C.c_void_p),
                         npar.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p),
                         stats.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(lib().b2emu_last_error().decode())
    keys = ("passes", "rounds", "dense", "factored", "fused_stores", "plain_passes")
    return st, dict(zip(keys, (int(x) for x in stats)))
