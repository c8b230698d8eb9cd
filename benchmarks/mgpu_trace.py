#!/usr/bin/env python
"""Multi-GPU timeline of one bench step (torchrun, one rank per GPU): every tile pass and every
exchange between CUDA events (b2sv_trace_begin/_end), printed per rank as totals and as a timeline.
Also sweeps the number of CTAs the exchange kernel may use (NVLink GB/s per direction vs SMs)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--sweep", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from bench import layer_circuit
    from pennylane_lightning_kokkos_b200 import dist as b2dist
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    n = args.qubits + g
    circ = layer_circuit(n, args.layers)
    oplist = ops.OpsStructKokkos_C128([c[0] for c in circ], [c[3] for c in circ], [c[1] for c in circ],
                                      [c[2] for c in circ])
    sv = b2dist.create_sharded_state(ops, n, np.complex128, local_rank)
    had = ops.OpsStructKokkos_C128(["Hadamard"] * n, [[] for _ in range(n)], [[w] for w in range(n)],
                                   [False] * n)
    sv.apply_ops(had)
    for _ in range(2):
        sv.apply_ops(oplist)
    sv.sync()
    dist.barrier()
    out = {"world": world, "qubits": n, "rank": rank}
    sv.trace_begin()
    for _ in range(args.steps):
        sv.apply_ops(oplist)
    tr = sv.trace_end()
    span = max(s + d for _, s, d in tr) - min(s for _, s, _ in tr)
    out["step_ms"] = span / args.steps
    out["pass_ms_per_step"] = sum(d for k, _, d in tr if k == 0) / args.steps
    out["passes_per_step"] = sum(1 for k, _, _ in tr if k == 0) / args.steps
    out["exchange_ms_per_step"] = sum(d for k, _, d in tr if k == 2) / args.steps
    out["exchanges_per_step"] = sum(1 for k, _, _ in tr if k == 2) / args.steps
    out["pass_ms"] = [round(d, 3) for k, _, d in tr if k == 0][:40]
    out["exchange_ms"] = [round(d, 3) for k, _, d in tr if k == 2][:20]
    out["gap_ms_per_step"] = out["step_ms"] - out["pass_ms_per_step"] - out["exchange_ms_per_step"]
    out["timeline_first_step"] = [(k, round(s, 2), round(d, 2)) for k, s, d in tr[: len(tr) // args.steps]]
    out["comm"] = sv.comm_stats()
    if args.sweep:
        # exchange bandwidth vs CTAs: PauliX on the global wire forces ... no -- use RX on wire 0 twice
        res = {}
        S = 16 * (1 << args.qubits)
        for ctas in [int(x) for x in args.sweep.split(",")]:
            os.environ["B2SV_EXCHANGE_CTAS"] = str(ctas)
            ts = []
            for rep in range(4):
                # a rotation on whichever wire currently sits on the top rank bit forces one exchange
                c0 = sv.comm_stats()
                sv.trace_begin()
                sv.apply(["RX"] * g, [[w] for w in sv_global_wires(sv, n, g)], [False] * g, [[0.3]] * g)
                t = sv.trace_end()
                c1 = sv.comm_stats()
                x = [d for k, _, d in t if k == 2]
                if x and rep > 0:
                    ts.append((sum(x), c1["swap_bytes_per_rank"] - c0["swap_bytes_per_rank"]))
            if ts:
                res[ctas] = {"ms": float(np.mean([a for a, _ in ts])),
                             "GBps_per_dir": float(np.mean([b / (a * 1e-3) / 1e9 for a, b in ts]))}
        os.environ.pop("B2SV_EXCHANGE_CTAS", None)
        out["exchange_sweep"] = res
    line = json.dumps(out)
    if rank == 0 or rank == world - 1:
        print(line, flush=True)
    if args.out and rank == 0:
        with open(args.out, "w") as f:
            f.write(line + "\n")
    del sv
    dist.barrier()
    dist.destroy_process_group()


def sv_global_wires(sv, n, g):
    """Wires whose qubits currently sit on rank bits (b2sv_layout); falls back to wires 0..g-1."""
    if hasattr(sv, "layout"):
        l2p = sv.layout()
        return [n - 1 - q for q in range(n) if l2p[q] >= n - g]
    return list(range(g))


if __name__ == "__main__":
    main()
