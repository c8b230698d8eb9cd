#!/usr/bin/env python
"""Randomised sweep of the scheduler through the pass emulator (CPU only, not collected by pytest): random and
layered circuits on 12-19 qubits, both table precisions, tile_low 3-6, every arithmetic budget, factored rounds
on/off, fused store on/off, against oracle/np_oracle.py. TEST INFRASTRUCTURE.

usage: python tests/fuzz_schedule.py [n_cases=600] [first_seed=2000]
       python tests/fuzz_schedule.py plan [n_cases=300] [first_seed=3000]   (the sharded-state planner, worlds 2-16,
                                                                             executed in NumPy as in test_shard_plan.py)
Last runs (round 2): 600 cases, 324 s: worst relative error 3.4e-15 (complex128 tables), 9.7e-7 (complex64), 0 failures;
plan: 300 cases, 21 s: worst absolute error 3.6e-16, 0 failures.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import emu  # noqa: E402
from cases import layered_circuit, random_circuit  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402


def main_plan(argv):
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as ops
    from test_shard_plan import run_plan, split
    n_cases = int(argv[0]) if argv else 300
    first = int(argv[1]) if len(argv) > 1 else 3000
    t0, worst, bad = time.time(), 0.0, 0
    for seed in range(first, first + n_cases):
        rng = np.random.default_rng(seed)
        world = int(rng.choice([2, 4, 8, 16]))
        n = int(rng.integers(world.bit_length() - 1 + 5, 13))
        if rng.integers(3) == 0:
            circ = layered_circuit(n, int(rng.integers(1, 5)), seed=seed)
        else:
            circ = random_circuit(n, int(rng.integers(5, 200)), seed=seed)
        plan = ops.OpsStructKokkos_C128(*split(circ)).plan_sharded(n, world, with_text=True)
        psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi /= np.linalg.norm(psi)
        e = np.max(np.abs(run_plan(plan["text"], n, psi) - npo.apply_ops(psi, n, circ)))
        worst = max(worst, e)
        if e > 1e-11:
            bad += 1
            print("FAIL seed", seed, "n", n, "world", world, "ops", len(circ), "err", e)
    print(f"plan: {n_cases} cases, worst absolute error {worst:.2e}, {bad} failures, {time.time() - t0:.0f} s")
    return 1 if bad else 0


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "plan":
        return main_plan(sys.argv[2:])
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    t0, worst, bad = time.time(), {False: 0.0, True: 0.0}, 0
    for seed in range(first, first + n_cases):
        rng = np.random.default_rng(seed)
        n, f32 = int(rng.integers(12, 20)), bool(rng.integers(2))
        if rng.integers(4) == 0:
            circ = layered_circuit(n, int(rng.integers(1, 5)), seed=seed)
        else:
            circ = random_circuit(n, int(rng.integers(10, 300)), seed=seed)
        psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi /= np.linalg.norm(psi)
        mh, low = int(rng.choice([0, 8, 12, 16, 24])), int(rng.choice([3, 4, 5, 6]))
        factor, sm = bool(rng.integers(2)), int(rng.integers(2))
        got, _ = emu.run(circ, n, psi, f32=f32, low=low, max_heavy=mh, factor=factor, store_mode=sm)
        want = npo.apply_ops(psi, n, circ)
        e = np.max(np.abs(got - want)) / np.max(np.abs(want))
        worst[f32] = max(worst[f32], e)
        if e > (2e-4 if f32 else 1e-11):
            bad += 1
            print("FAIL seed", seed, "n", n, "f32", f32, "ops", len(circ), "max_heavy", mh, "low", low,
                  "factor", factor, "store_mode", sm, "rel err", e)
    print(f"{n_cases} cases, worst relative error c128 {worst[False]:.2e} / c64 {worst[True]:.2e}, "
          f"{bad} failures, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
