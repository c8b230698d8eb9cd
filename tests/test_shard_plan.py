"""CPU tests of the sharded-state planner (csrc/shard_plan.cpp through b2sv_plan_sharded).

The plan -- runs of primitives in physical bits + k-bit exchanges -- is executed in NumPy on a full
2^n vector indexed by PHYSICAL bits (an exchange swaps index-bit pairs, exactly what the all-to-all
between the shards does) and compared with the NumPy oracle on the logical circuit."""
import numpy as np
import pytest

from cases import layered_circuit, random_circuit
from oracle import np_oracle as npo


@pytest.fixture(scope="module")
def ops():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    return m


def split(circ):
    return ([c[0] for c in circ], [c[3] for c in circ], [c[1] for c in circ], [c[2] for c in circ])


def run_plan(text, n, psi):
    """psi: state indexed by physical bits (identity layout at the start)."""
    psi = psi.copy()
    idx = np.arange(1 << n, dtype=np.int64)
    l2p = None
    for line in text.strip().split("\n"):
        f = line.split()
        if f[0] == "RUN":
            continue
        if f[0] == "EXCH":
            v = [int(x) for x in f[1:]]
            perm = idx.copy()
            for gp, lp in zip(v[0::2], v[1::2]):
                bg, bl = (perm >> gp) & 1, (perm >> lp) & 1
                perm = perm & ~((1 << gp) | (1 << lp)) | (bl << gp) | (bg << lp)
            new = np.empty_like(psi)
            new[perm] = psi  # the amplitude at physical index i moves to i with the bit pairs swapped
            psi = new
        elif f[0] == "C1Q":
            t, cm, cv = int(f[1]), int(f[2]), int(f[3])
            m = np.array([float(x) for x in f[4:12]]).view(np.complex128).reshape(2, 2)
            sel0 = ((idx & cm) == cv) & (((idx >> t) & 1) == 0)
            i0 = idx[sel0]
            i1 = i0 | (1 << t)
            a0, a1 = psi[i0], psi[i1]
            psi[i0] = m[0, 0] * a0 + m[0, 1] * a1
            psi[i1] = m[1, 0] * a0 + m[1, 1] * a1
        elif f[0] == "DIAG":
            pm, cm, cv = int(f[1]), int(f[2]), int(f[3])
            p = np.array([float(x) for x in f[4:8]]).view(np.complex128)
            par = np.zeros(1 << n, dtype=np.int64)
            mm = pm
            while mm:
                par ^= (idx >> ((mm & -mm).bit_length() - 1)) & 1
                mm &= mm - 1
            sel = (idx & cm) == cv
            psi[sel] *= np.where(par[sel] == 1, p[1], p[0])
        elif f[0] == "MATK":
            k = int(f[1])
            bits = [int(x) for x in f[2:2 + k]]
            mat = np.array([float(x) for x in f[2 + k:]]).view(np.complex128).reshape(1 << k, 1 << k)
            psi = npo.apply_matrix(psi, n, mat, [n - 1 - b for b in bits])
        elif f[0] == "L2P":
            l2p = [int(x) for x in f[1:]]
    # back to logical order: logical bit q sits at physical bit l2p[q]
    phys = np.zeros(1 << n, dtype=np.int64)
    for q in range(n):
        phys |= ((idx >> q) & 1) << l2p[q]
    return psi[phys]


@pytest.mark.parametrize("n,world", [(8, 2), (9, 4), (10, 8)])
@pytest.mark.parametrize("kind", ["layers", "random"])
def test_plan_is_correct(ops, n, world, kind):
    circ = layered_circuit(n, 3, seed=4) if kind == "layers" else random_circuit(n, 150, seed=9)
    names, params, wires, invs = split(circ)
    o = ops.OpsStructKokkos_C128(names, params, wires, invs)
    plan = o.plan_sharded(n, world, with_text=True)
    rng = np.random.default_rng(1)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    got = run_plan(plan["text"], n, psi0)
    want = npo.apply_ops(psi0, n, circ)
    assert np.max(np.abs(got - want)) < 1e-12
    if kind == "layers":
        assert plan["exchanges"] >= 1


def test_plan_config5_shape(ops):
    """BASELINE config 5 (weak scaling 33 -> 36 qubits, 33 local): about one exchange and one layer's
    worth of tile passes per layer, every global qubit of a layer moved in ONE all-to-all."""
    for world, n in ((2, 34), (4, 35), (8, 36)):
        g = world.bit_length() - 1
        circ = layered_circuit(n, 2, seed=42)
        names, params, wires, invs = split(circ)
        o = ops.OpsStructKokkos_C128(names, params, wires, invs)
        plan = o.plan_sharded(n, world)
        assert plan["exchanges"] <= 3
        assert plan["exchanged_bits"] <= 3 * g
        assert plan["passes"] <= 16
        shard = 16 * (1 << (n - g))
        assert plan["bytes_per_rank"] <= 3 * (1 - 0.5 ** g) * shard + 1
