"""Multi-GPU parity worker, launched by torchrun (one rank per GPU) from tests/test_multi_gpu.py.

Every rank applies the same random circuit to a sharded state (rank = top index bits); the shards
are gathered and compared with (a) the NumPy oracle, (b) the single-GPU engine on rank 0, and the
all-reduced expectation values / adjoint Jacobian are compared with the oracle's.  Prints one
"MGPU_OK ..." line on rank 0 when everything agrees.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    from cases import layered_circuit, random_circuit
    from oracle import np_oracle as npo
    from pennylane_lightning_kokkos_b200 import dist as b2dist
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12 + 2 * g + 1
    worst = 0.0
    sliced_runs = 0
    # every case twice: whole-shard execution, and the pipelined path (shard cut into 4 slices, exchanges
    # on a second stream) forced on these small states through B2SV_PIPE_MIN_SUB
    for mode, dtype, tol in (("plain", np.complex128, 1e-12), ("plain", np.complex64, 1e-5),
                             ("sliced", np.complex128, 1e-12), ("sliced", np.complex64, 1e-5)):
        if mode == "sliced":
            os.environ["B2SV_PIPE_MIN_SUB"] = "0"
            n_eff = 18 + g + (0 if dtype == np.complex128 else 1)
        else:
            os.environ.pop("B2SV_PIPE_MIN_SUB", None)
            n_eff = n if dtype == np.complex128 else n + 1
        for case, circ in (("layers", layered_circuit(n_eff, 3 if mode == "sliced" else 2, seed=3)),
                           ("random", random_circuit(n_eff, 120, seed=5))):
            sv = b2dist.create_sharded_state(ops, n_eff, dtype, local_rank)
            sv.reset_stats()
            sv.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
                     [c[3] for c in circ])
            stats = sv.comm_stats()
            st = sv.stats()
            if st["launches"] > st["sweeps"]:
                sliced_runs += 1
            # expectation values before normalising the layout (exercise the remapped reductions)
            ez = [sv.ExpectationValue("PauliZ", [w], [], np.zeros(0)) for w in (0, g, n_eff - 1)]
            ex = [sv.ExpectationValue("PauliX", [w], [], np.zeros(0)) for w in (0, n_eff - 1)]
            nloc = 1 << (n_eff - g)
            mine = np.zeros(nloc, dtype=dtype)
            sv.DeviceToHost(mine)
            t = torch.from_numpy(mine.view(np.float64 if dtype == np.complex128 else np.float32)).cuda()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            full = np.concatenate([p.cpu().numpy() for p in parts]).view(dtype)
            psi0 = np.zeros(1 << n_eff, dtype=complex)
            psi0[0] = 1
            want = npo.apply_ops(psi0, n_eff, circ)
            err = float(np.max(np.abs(full - want)) / np.max(np.abs(want)))
            ez_want = [npo.expval(want, n_eff, ("named", "PauliZ", [w])) for w in (0, g, n_eff - 1)]
            ex_want = [npo.expval(want, n_eff, ("named", "PauliX", [w])) for w in (0, n_eff - 1)]
            e2 = max(abs(a - b) for a, b in zip(ez + ex, ez_want + ex_want))
            worst = max(worst, err / tol, e2 / tol)
            if rank == 0:
                print(f"  {mode} {case} {np.dtype(dtype).name} n={n_eff} world={world}: amp err {err:.2e} "
                      f"expval err {e2:.2e} swaps {stats['swaps']} path {stats['path']}", flush=True)
            assert err < tol and e2 < tol, (case, dtype, err, e2)
            # a Pauli-sum Hamiltonian: applied by one kernel that reads the partner amplitudes of terms with
            # X / Y factors on global qubits from the peers' shards (expval = <psi|H psi>, var uses it twice)
            if dtype == np.complex128:
                from cases import random_pauli_hamiltonian
                sfx = "C128"
                ham = random_pauli_hamiltonian(n_eff, 16, seed=11)
                ham[0] = (0.7, [("PauliX", 0), ("PauliY", n_eff - 1)])  # global X, local Y
                ham[1] = (-0.4, [("PauliZ", 0), ("PauliX", g)])        # global Z, first local X
                tobs = []
                for _, word in ham:
                    fac = [getattr(ops, "NamedObsKokkos_" + sfx)(l, [w]) for l, w in word]
                    tobs.append(fac[0] if len(fac) == 1 else getattr(ops, "TensorProdObsKokkos_" + sfx)(fac))
                Hobs = getattr(ops, "HamiltonianKokkos_" + sfx)(np.array([c for c, _ in ham]), tobs)
                hnp = ("hamiltonian", [c for c, _ in ham],
                       [("tensor", [("named", l, [w]) for l, w in word]) for _, word in ham])
                eh, eh_want = sv.expval(Hobs), npo.expval(want, n_eff, hnp)
                vh, vh_want = sv.var(Hobs), npo.var(want, n_eff, hnp)
                assert abs(eh - eh_want) < 1e-12 and abs(vh - vh_want) < 1e-11, (eh, eh_want, vh, vh_want)
                if mode == "plain":
                    # the same operator as CSR: every rank streams its rows, columns on other ranks are
                    # read from the peers' shards
                    import scipy.sparse as sp
                    dim = 1 << n_eff
                    idx = np.arange(dim, dtype=np.int64)
                    mat = sp.csr_matrix((dim, dim), dtype=np.complex128)
                    for c, word in ham:
                        x = z = ny = 0
                        for l, w in word:
                            b = 1 << (n_eff - 1 - w)
                            x |= b if l in ("PauliX", "PauliY") else 0
                            z |= b if l in ("PauliZ", "PauliY") else 0
                            ny += l == "PauliY"
                        par = np.zeros(dim, dtype=np.int64)
                        zz = z
                        while zz:
                            par ^= (idx >> ((zz & -zz).bit_length() - 1)) & 1
                            zz &= zz - 1
                        mat = mat + sp.csr_matrix((c * (1j ** ny) * (1 - 2 * par), (idx ^ x, idx)), shape=(dim, dim))
                    mat = mat.tocsr()
                    mat.sort_indices()
                    ec = sv.ExpectationValue(mat.data, mat.indices.astype(np.uint64), mat.indptr.astype(np.uint64))
                    assert abs(ec - eh_want) < 1e-12, (ec, eh_want)
            # marginal probabilities over global + local wires (all-reduced histograms) and sampling
            # (one shared random stream; each shot resolved by the rank that owns its interval)
            for pw in ([0, g, n_eff - 1], [1, 2], list(range(min(n_eff, 10)))):
                pw = sorted(set(pw))
                got_p = sv.probs(pw)
                want_p = npo.probs(want, n_eff, pw)
                ep = float(np.max(np.abs(got_p - want_p)))
                assert ep < (1e-12 if dtype == np.complex128 else 2e-6), (case, dtype, pw, ep)
            shots = 4000
            smp = sv.GenerateSamples(n_eff, shots)
            assert smp.shape == (shots, n_eff) and smp.max() <= 1
            t_s = torch.from_numpy(smp.astype(np.int64)).cuda()
            t_0 = t_s.clone()
            dist.broadcast(t_0, src=0)
            assert bool((t_s == t_0).all()), "ranks disagree on the samples"
            freq = smp[:, [0, n_eff - 1]].mean(axis=0)  # P(bit = 1) of a global and a local wire
            p1 = [npo.probs(want, n_eff, [w])[1] for w in (0, n_eff - 1)]
            assert max(abs(a - b) for a, b in zip(freq, p1)) < 0.05, (freq, p1)
            if rank == 0 and dtype == np.complex128:
                single = ops.LightningKokkos_C128(n_eff)
                single.apply([c[0] for c in circ], [c[1] for c in circ], [c[2] for c in circ],
                             [c[3] for c in circ])
                ref1 = np.zeros(1 << n_eff, dtype=dtype)
                single.DeviceToHost(ref1)
                assert np.max(np.abs(ref1 - full)) < 1e-13
                del single
            del sv
            dist.barrier()
    os.environ.pop("B2SV_PIPE_MIN_SUB", None)
    assert sliced_runs >= 2, f"the pipelined path never ran ({sliced_runs})"
    # adjoint Jacobian on a sharded state (all-reduced inner products)
    n_a = 12 + 2 * g
    circ = [c for c in layered_circuit(n_a, 1, seed=7)]
    names, wires = [c[0] for c in circ], [c[1] for c in circ]
    invs, params = [c[2] for c in circ], [c[3] for c in circ]
    sv = b2dist.create_sharded_state(ops, n_a, np.complex128, local_rank)
    sv.apply(names, wires, invs, params)
    obs = [ops.NamedObsKokkos_C128("PauliZ", [0]), ops.NamedObsKokkos_C128("PauliX", [n_a - 1])]
    adj = ops.AdjointJacobianKokkos_C128()
    oplist = adj.create_ops_list(names, [np.array(p) for p in params], wires, invs,
                                 [np.zeros(0, dtype=complex) for _ in names])
    n_par = sum(1 for p in params if len(p))
    tp = list(range(n_par))
    jac = adj.adjoint_jacobian(sv, obs, oplist, tp)
    psi0 = np.zeros(1 << n_a, dtype=complex)
    psi0[0] = 1
    final = npo.apply_ops(psi0, n_a, circ)
    jac_want = npo.adjoint_jacobian(final, n_a, [("named", "PauliZ", [0]),
                                                 ("named", "PauliX", [n_a - 1])], circ, tp)
    ej = float(np.max(np.abs(jac - jac_want)) / max(np.max(np.abs(jac_want)), 1e-300))
    assert ej < 1e-12, ej
    del sv
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} worst_err_over_tol={worst:.3f} adjoint_rel_err={ej:.2e} "
              f"sliced_runs={sliced_runs}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
