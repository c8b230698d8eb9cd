"""CPU tests of the host logic that decides what the tile executor does: named gates are lowered,
scheduled into passes / rounds (free permutations, dense + factored rounds, fused stores) and the
resulting descriptors are re-executed by the scalar pass emulator (tests/emu.py), then compared
with the NumPy oracle. No GPU, no product library involved."""
import numpy as np
import pytest

import emu
from cases import gate_cases, layered_circuit, random_circuit, random_state, sel_circuit
from oracle import np_oracle as npo


def check(circ, n, f32=False, tol=None, **kw):
    psi = random_state(n, seed=len(circ) + n)
    got, stats = emu.run(circ, n, psi, f32=f32, **kw)
    want = npo.apply_ops(psi, n, circ)
    err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    assert err < (tol or (2e-5 if f32 else 1e-12)), (err, stats)
    return stats


@pytest.mark.parametrize("factor", [True, False])
def test_layered_circuit_factored_rounds(factor):
    n = 14
    circ = layered_circuit(n, 3, seed=42)
    st = check(circ, n, factor=factor)
    assert (st["factored"] > 0) == factor
    # every CNOT and every X split off an anti-diagonal-dominant 2x2 is absorbed: no generic rounds
    assert st["rounds"] == st["dense"] + st["factored"]


@pytest.mark.parametrize("max_heavy", [0, 1, 4, 12, 16])
def test_layered_circuit_pass_budgets(max_heavy):
    n = 13
    check(layered_circuit(n, 2, seed=7), n, max_heavy=max_heavy)


def test_layered_circuit_c64_tables_in_float():
    n = 14
    st = check(layered_circuit(n, 2, seed=3), n, f32=True)
    assert st["factored"] > 0


@pytest.mark.parametrize("store_mode", [0, 1])
def test_store_modes(store_mode):
    n = 14
    st = check(layered_circuit(n, 3, seed=21), n, store_mode=store_mode)
    check(random_circuit(n, 200, seed=store_mode), n, store_mode=store_mode)
    assert (st["fused_stores"] > 0) == (store_mode == 1)


@pytest.mark.parametrize("B,low", [(11, 5), (12, 4), (12, 6)])
def test_tile_geometries(B, low):
    n = 13
    check(layered_circuit(n, 2, seed=11), n, B=B, low=low)
    check(random_circuit(n, 120, seed=B * 10 + low), n, B=B, low=low)


def test_antidiagonal_dominant_gates_split_into_x():
    # RX/RY near pi have a vanishing diagonal: X * (X M) keeps the shears bounded
    n = 13
    circ = []
    for w in range(n):
        circ.append(("RX", [w], False, [np.pi - 1e-3 * w]))
        circ.append(("RY", [(w + 3) % n], False, [np.pi + 1e-9 * w]))
        circ.append(("PauliY", [(w + 5) % n], False, []))
        circ.append(("Hadamard", [(w + 7) % n], False, []))
    st = check(circ, n)
    assert st["factored"] > 0


@pytest.mark.parametrize("seed", range(6))
def test_random_circuits_every_gate(seed):
    n = 13
    check(random_circuit(n, 250, seed=seed), n)


def test_random_circuits_c64():
    n = 14
    check(random_circuit(n, 150, seed=99), n, f32=True, tol=5e-5)


def test_every_gate_pattern():
    n = 13
    for case in gate_cases(n, seed=5, per_gate=3):
        check([case], n)


def test_sel_circuit():
    check(sel_circuit(13, 3), 13)


def test_long_circuit_bounded_scan():
    """Scheduling uses a bounded look-ahead window and gives up on a pass after 512 misses in a row;
    a 3000-gate circuit (more pending ops than either bound) must still come out exact."""
    n = 13
    circ = random_circuit(n, 3000, seed=123)
    st = check(circ, n, max_heavy=0, tol=5e-12)
    assert st["passes"] < 1200


@pytest.mark.parametrize("seed", range(4))
def test_random_geometry_and_budget(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(12, 15))
    f32 = bool(rng.integers(2))
    B = int(rng.integers(11, 13)) + (1 if f32 else 0)
    low = int(rng.integers(2, 7))
    mh = int(rng.choice([0, 3, 5, 9, 24]))
    circ = random_circuit(n, 200, seed=seed) + layered_circuit(n, 1, seed=seed)
    check(circ, n, f32=f32, B=B, low=low, max_heavy=mh, factor=bool(rng.integers(2)),
          store_mode=int(rng.integers(2)), tol=1e-4 if f32 else None)


@pytest.mark.parametrize("f32", [False, True])
def test_reference_param_gate_literals_through_the_scheduler(f32):
    """The reference's own literal in/out vectors (tests/golden/ref_param_literals.json) through
    lower -> schedule -> emulated passes, both table precisions."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_param_literals.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        ini = np.array([complex(a, b) for a, b in c["ini"]])
        want = np.array([complex(a, b) for a, b in c["expected"]])
        n = int(np.log2(ini.size))
        got, _ = emu.run([(c["gate"], c["wires"], c["inverse"], c["params"])], n, ini, f32=f32)
        assert np.max(np.abs(got - want)) < 2e-6, c["gate"]  # the literals carry ~7 digits


def test_plain_layout_passes_for_bulk_loads(monkeypatch):
    """With bulk tile loads allowed, passes whose rounds keep the lowest index bits out of the register
    set are laid out in plain order (filled by cp.async.bulk on the GPU); results are unchanged."""
    import emu
    monkeypatch.setenv("B2EMU_BULK", "1")
    n = 16
    for f32 in (False, True):
        circ = layered_circuit(n, 3, seed=17)
        psi0 = np.zeros(1 << n, dtype=complex)
        psi0[0] = 1
        got, st = emu.run(circ, n, psi0, f32=f32)
        want = npo.apply_ops(psi0, n, circ)
        assert np.max(np.abs(got - want)) < (2e-5 if f32 else 1e-12)
        assert st["plain_passes"] >= 1 and st["plain_passes"] < st["passes"]
    circ = random_circuit(n, 200, seed=23)
    got, st = emu.run(circ, n, psi0)
    assert np.max(np.abs(got - npo.apply_ops(psi0, n, circ))) < 1e-12


def test_tile_swizzle_properties():
    """phys_slot (schedule.hpp) = the tile kernel's phys<B,SW,SH>: a GF(2)-linear bijection of the tile; any 8
    consecutive 16-byte units land in 8 distinct 16-byte bank groups of a 128-byte wavefront whatever higher bits
    are set; complex64 (SH = 1) keeps the two amplitudes of a 16-byte unit together, which is what lets the load
    warps copy 16 bytes at a time for both types."""
    import ctypes as C
    import emu
    f = emu.lib().b2emu_phys_slot
    f.restype = C.c_uint32
    f.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int]
    for B, SW, SH in ((12, 3, 0), (13, 4, 1)):
        n = 1 << B
        slots = np.array([f(i, B, SW, SH) for i in range(n)], dtype=np.int64)
        assert sorted(slots.tolist()) == list(range(n))                     # bijection
        rng = np.random.default_rng(B)
        for _ in range(200):                                                # linear over GF(2)
            a, b = (int(x) for x in rng.integers(n, size=2))
            assert slots[a ^ b] == slots[a] ^ slots[b]
        unit = slots >> SH                                                  # 16-byte unit index of each slot
        for base in rng.integers(n >> (SH + 3), size=64):                   # 8 consecutive units, any base
            first = int(base) << (SH + 3)
            groups = {int(unit[first + (u << SH)]) & 7 for u in range(8)}
            assert len(groups) == 8
        if SH:
            even = np.arange(0, n, 2)
            assert np.all(slots[even] % 2 == 0) and np.all(slots[even + 1] == slots[even] + 1)
