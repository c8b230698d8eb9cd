"""pennylane_lightning_kokkos_b200 -- a B200-native state-vector engine behind the
`lightning.kokkos` API surface (reference: PennyLane-Lightning-Kokkos).

Layout:
  csrc/                          CUDA kernels (sm_100a) + C++ host + the C ABI -> libb2sv.so
  _lib.py                        ctypes loader (fails loudly if the library is missing)
  lightning_kokkos_qubit_ops.py  ctypes mirror of the reference's pybind11 module (same class/method names)
  lightning_kokkos_qubit_ops_pyb compiled pybind11 module with the same surface (csrc/pybind/bindings.cpp)
  lightning_kokkos.py            PennyLane-free mirror of the reference's device class
  dist.py                        torch.distributed plumbing for sharded states
"""
from ._lib import PLException, backend_info, device_count  # noqa: F401

__version__ = "0.1.0"
