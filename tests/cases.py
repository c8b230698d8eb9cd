"""Shared test-case generators: gate catalogue, wire patterns, random states and circuits."""
import itertools

import numpy as np

# name -> (number of wires, number of params); 0 wires = variable (MultiRZ)
GATES = {
    "PauliX": (1, 0), "PauliY": (1, 0), "PauliZ": (1, 0), "Hadamard": (1, 0), "S": (1, 0),
    "T": (1, 0), "PhaseShift": (1, 1), "RX": (1, 1), "RY": (1, 1), "RZ": (1, 1), "Rot": (1, 3),
    "CNOT": (2, 0), "CY": (2, 0), "CZ": (2, 0), "SWAP": (2, 0), "ControlledPhaseShift": (2, 1),
    "CRX": (2, 1), "CRY": (2, 1), "CRZ": (2, 1), "CRot": (2, 3), "IsingXX": (2, 1),
    "IsingXY": (2, 1), "IsingYY": (2, 1), "IsingZZ": (2, 1), "SingleExcitation": (2, 1),
    "SingleExcitationMinus": (2, 1), "SingleExcitationPlus": (2, 1), "CSWAP": (3, 0),
    "Toffoli": (3, 0), "DoubleExcitation": (4, 1), "DoubleExcitationMinus": (4, 1),
    "DoubleExcitationPlus": (4, 1), "MultiRZ": (0, 1),
}
GENERATORS = ["RX", "RY", "RZ", "PhaseShift", "ControlledPhaseShift", "CRX", "CRY", "CRZ", "IsingXX",
              "IsingXY", "IsingYY", "IsingZZ", "SingleExcitation", "SingleExcitationMinus",
              "SingleExcitationPlus", "DoubleExcitation", "DoubleExcitationMinus",
              "DoubleExcitationPlus", "MultiRZ"]


def random_state(n, seed, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def wire_patterns(n, k, rng, count):
    """`count` random ordered k-subsets of range(n), always including low/high extremes."""
    pats = []
    base = list(range(n))
    if k <= n:
        pats.append(base[:k])
        pats.append(base[-k:][::-1])
        pats.append(([0] + base[-(k - 1):]) if k > 1 else [n - 1])
    while len(pats) < count:
        pats.append([int(x) for x in rng.choice(n, size=k, replace=False)])
    out, seen = [], set()
    for p in pats:
        if tuple(p) not in seen and len(set(p)) == k:
            seen.add(tuple(p))
            out.append(p)
    return out


def gate_cases(n, seed=0, per_gate=4):
    """(name, wires, inverse, params) covering every gate x wire pattern x inverse."""
    rng = np.random.default_rng(seed)
    cases = []
    for name, (nw, npar) in GATES.items():
        ks = [nw] if nw else [1, 2, 3]
        for k in ks:
            if k > n:
                continue
            for wires in wire_patterns(n, k, rng, per_gate):
                for inv in (False, True):
                    params = [float(x) for x in rng.uniform(-np.pi, np.pi, size=npar)]
                    cases.append((name, wires, inv, params))
    return cases


def random_circuit(n, depth, seed, names=None):
    rng = np.random.default_rng(seed)
    names = names or list(GATES)
    ops = []
    for _ in range(depth):
        name = names[int(rng.integers(len(names)))]
        nw, npar = GATES[name]
        k = nw if nw else int(rng.integers(1, min(n, 4) + 1))
        if k > n:
            continue
        wires = [int(x) for x in rng.choice(n, size=k, replace=False)]
        params = [float(x) for x in rng.uniform(-np.pi, np.pi, size=npar)]
        ops.append((name, wires, bool(rng.integers(2)), params))
    return ops


def layered_circuit(n, layers, seed):
    """BASELINE config 2/5 shape: RX,RY,RZ on every wire + CNOT ring, per layer."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(layers):
        for w in range(n):
            for g in ("RX", "RY", "RZ"):
                ops.append((g, [w], False, [float(rng.uniform(0, 2 * np.pi))]))
        for w in range(n):
            ops.append(("CNOT", [w, (w + 1) % n], False, []))
    return ops


def sel_circuit(n, layers, seed=42):
    """BASELINE config 1: StronglyEntanglingLayers (Rot per wire + CNOT ring of range r)."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(0, 2 * np.pi, size=(layers, n, 3))
    ops = []
    for l in range(layers):
        for i in range(n):
            ops.append(("Rot", [i], False, [float(x) for x in w[l, i]]))
        r = (l % (n - 1)) + 1 if n > 1 else 0
        if n > 1:
            for i in range(n):
                ops.append(("CNOT", [i, (i + r) % n], False, []))
    return ops


def random_pauli_hamiltonian(n, terms, seed, max_weight=4):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(terms):
        k = int(rng.integers(1, min(max_weight, n) + 1))
        wires = [int(x) for x in rng.choice(n, size=k, replace=False)]
        letters = [["PauliX", "PauliY", "PauliZ"][int(rng.integers(3))] for _ in wires]
        out.append((float(rng.uniform(-1, 1)), list(zip(letters, wires))))
    return out
