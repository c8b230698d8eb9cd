#!/usr/bin/env python
"""Developer experiment (torchrun, 2 ranks): cost of one global<->local qubit swap on a sharded
30-qubit-per-GPU c128 state, next to the plain NVLink peer-copy bandwidth of the box."""
import json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops, dist as b2dist
rank = dist.get_rank()
# plain peer copy (torch uses cudaMemcpyPeerAsync): 4 GiB, both directions at once
a = torch.empty(1 << 32, dtype=torch.uint8, device=f"cuda:{lr}")
other = torch.empty(1 << 32, dtype=torch.uint8, device=f"cuda:{1 - lr}")
torch.cuda.synchronize(); dist.barrier()
for _ in range(2):
    t0 = time.perf_counter(); other.copy_(a); torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0: print(json.dumps({"peer_copy_gbs_per_direction_both_active": (1 << 32) / dt / 1e9}), flush=True)
del a, other; torch.cuda.empty_cache(); dist.barrier()
n = int(os.environ.get("QPG", "30")) + 1
sv = b2dist.create_sharded_state(ops, n, np.complex128, lr)
for w in range(n):
    sv.Hadamard([w], False, [])
sv.sync(); dist.barrier()
for wire_pair in [(0, 5), (0, 12), (0, 25), (0, n - 1)]:
    # RX on the global wire forces a swap-in; RX on the evicted wire forces the next one
    times = []
    for rep in range(6):
        s0 = sv.comm_stats()["swaps"]
        t0 = time.perf_counter()
        sv.RX([wire_pair[rep % 2]], False, [0.1])
        sv.sync(); dt = time.perf_counter() - t0
        times.append((dt, sv.comm_stats()["swaps"] - s0))
    if rank == 0: print(json.dumps({"wires": wire_pair, "ms_and_swaps": [(round(t * 1e3, 2), s) for t, s in times]}), flush=True)
del sv
dist.barrier(); dist.destroy_process_group()
