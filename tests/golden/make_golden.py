"""Generates the golden fixtures from the COMPILED REFERENCE (oracle/_ref). Run in the build
container (needs /root/reference to build oracle/_ref):  python tests/golden/make_golden.py
The fixtures travel with the repo; the GPU box only reads them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import GATES, gate_cases, random_pauli_hamiltonian, random_state  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    ref.build()
    n = 5
    st = random_state(n, 123)
    sv = ref.RefStateVector(n)
    cases = gate_cases(n, seed=7, per_gate=3)
    outs = []
    for name, wires, inv, params in cases:
        sv.h2d(st)
        sv.apply(name, wires, inv, params)
        outs.append(sv.d2h())
    np.savez_compressed(os.path.join(HERE, "gates_n5.npz"), state=st, out=np.array(outs),
                        cases=np.array(cases, dtype=object))

    n = 6
    rng = np.random.default_rng(99)
    par = [g for g, (nw, k) in GATES.items() if k == 1]
    circ = []
    for g in par + par[:6]:
        nw = GATES[g][0] or 3
        wires = [int(x) for x in rng.choice(n, size=nw, replace=False)]
        circ.append((g, wires, bool(rng.integers(2)), [float(rng.uniform(-1, 1))]))
        circ.append(("CNOT", [int(x) for x in rng.choice(n, size=2, replace=False)], False, []))
    terms = random_pauli_hamiltonian(n, 8, seed=5)
    robs = []
    for _, word in terms:
        rf = [ref.RefObs.named(nm, [w]) for nm, w in word]
        robs.append(rf[0] if len(rf) == 1 else ref.RefObs.tensor(rf))
    ham = ref.RefObs.hamiltonian([c for c, _ in terms], robs)
    sv = ref.RefStateVector(n)
    sv.apply_ops(circ)
    n_par = sum(1 for c in circ if c[3])
    tp = sorted(int(x) for x in rng.choice(n_par, size=n_par - 5, replace=False))
    jac = sv.adjoint_jacobian([ham], circ, tp)
    np.savez_compressed(os.path.join(HERE, "adjoint_n6.npz"), circ=np.array(circ, dtype=object),
                        terms=np.array(terms, dtype=object), tp=np.array(tp), jac=jac,
                        state=sv.d2h(), expval=sv.expval_obs(ham))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
