"""torch.distributed plumbing for sharded state vectors (one process per GPU).

The engine shards a 2^n state over P = 2^g GPUs of one box: rank r holds the amplitudes whose top g
index bits equal r (wires 0..g-1 are "global").  Data moves only when a non-diagonal gate targets a
global wire: that wire is swapped with a shard-local one by an in-place NVLink peer-memory kernel
(csrc/kernels.cu k_peer_swap; NCCL send/recv through a staging buffer as the second path), and
expectation values / norms / inner products are all-reduced with NCCL.  The reference has no
multi-device code at all (SURVEY.md section 2 "Parallelism strategies"); this module only does the
rendezvous: it gets the 128-byte NCCL unique id from rank 0 to everybody over the process group the
launcher (torchrun) already set up -- gloo or nccl, so the same code runs in CPU-only tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def nccl_unique_id() -> bytes:
    """128 opaque bytes from ncclGetUniqueId (b2sv_comm_unique_id); call on rank 0 only."""
    from ._lib import check, lib
    buf = C.create_string_buffer(128)
    check(lib.b2sv_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast ``nbytes`` bytes from rank ``src`` over the default process group."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def shard_geometry(num_qubits: int, world: int) -> dict:
    """Which wires are global and how big each shard is (pure host arithmetic)."""
    g = world.bit_length() - 1
    if world < 1 or (1 << g) != world:
        raise ValueError("the number of ranks must be a power of two")
    if num_qubits - g < 1:
        raise ValueError("too few qubits for this many ranks")
    return {"global_wires": list(range(g)), "local_qubits": num_qubits - g,
            "amplitudes_per_rank": 1 << (num_qubits - g)}


def local_slice(full_state: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The contiguous slice of a full state vector that rank ``rank`` owns (top bits = rank)."""
    n = full_state.size // world
    return full_state[rank * n:(rank + 1) * n]


def create_sharded_state(ops_module, num_qubits: int, dtype, device_id: int):
    """Collective: build a sharded LightningKokkos_C64/C128 on every rank of the default group."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    shard_geometry(num_qubits, world)
    uid = nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    cls = ops_module.LightningKokkos_C128 if np.dtype(dtype) == np.complex128 else ops_module.LightningKokkos_C64
    return cls.sharded(num_qubits, device_id, rank, world, uid)
