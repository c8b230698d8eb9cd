"""Multi-GPU tests.

* CPU (`-m "not gpu"`): the host-side plumbing of sharded states over a world-size-2 gloo group --
  shard geometry, unique-id style byte broadcast, shard slicing.
* GPU (`-m gpu`): torchrun with 2 ranks (when the box has >= 2 GPUs) runs tests/mgpu_worker.py:
  sharded circuits vs the NumPy oracle and the single-GPU engine, all-reduced expectation values,
  and an adjoint Jacobian on a sharded state.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from pennylane_lightning_kokkos_b200 import dist as b2dist

        payload = bytes(range(128)) if rank == 0 else None
        got = b2dist.broadcast_bytes(payload, 128, src=0)
        geo = b2dist.shard_geometry(14, world)
        full = np.arange(1 << 6, dtype=np.complex128)
        mine = b2dist.local_slice(full, rank, world)
        q.put((rank, got == bytes(range(128)), geo, mine[0].real, mine.size))
    finally:
        dist.destroy_process_group()


def test_sharded_plumbing_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, geo, first, size in res:
        assert ok
        assert geo == {"global_wires": [0], "local_qubits": 13, "amplitudes_per_rank": 1 << 13}
        assert size == 32 and first == rank * 32  # rank = top index bit


def test_shard_geometry_errors():
    sys.path.insert(0, ROOT)
    from pennylane_lightning_kokkos_b200 import dist as b2dist

    with pytest.raises(ValueError):
        b2dist.shard_geometry(10, 3)
    with pytest.raises(ValueError):
        b2dist.shard_geometry(2, 4)
    assert b2dist.shard_geometry(36, 8)["local_qubits"] == 33


@pytest.mark.gpu
def test_sharded_state_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29600 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK" in out.stdout


@pytest.mark.gpu
def test_two_devices_in_one_process():
    """Function attributes (dynamic shared memory opt-in) and the SM count are per device: a process
    that holds states on two GPUs must be able to run the tile executor and the transition-sum kernel
    on both."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import layered_circuit
    from oracle import np_oracle as npo
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

    n = 14
    circ = layered_circuit(n, 2, seed=5)
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[0] = 1
    want = npo.apply_ops(psi0, n, circ)
    tp = list(range(sum(1 for p in params if p)))
    jac_want = npo.adjoint_jacobian(want, n, [("named", "PauliZ", [0])], circ, tp)
    for dev in (0, 1, 0):
        sv = ops.LightningKokkos_C128(n, ops.InitializationSettings().set_device_id(dev))
        sv.apply(names, wires, invs, params)
        got = np.zeros(1 << n, dtype=np.complex128)
        sv.DeviceToHost(got)
        assert np.max(np.abs(got - want)) < 1e-12
        adj = ops.AdjointJacobianKokkos_C128()
        ol = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs,
                                 [np.zeros(0, dtype=complex) for _ in names])
        jac = adj.adjoint_jacobian(sv, [ops.NamedObsKokkos_C128("PauliZ", [0])], ol, tp)
        assert np.max(np.abs(jac - jac_want)) < 1e-12
