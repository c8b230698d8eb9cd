#!/usr/bin/env python
"""Extracts the literal input / expected state vectors of the reference's own parametric-gate tests
(reference src/tests/Test_StateVectorKokkos_Param.cpp: blocks `std::vector<cp_t> ini_st{...}` +
`std::vector<cp_t> expected{...}` followed by `kokkos_sv.apply<Gate>({wires}, inverse, {params})`)
and the sparse-table style of the Ising/MultiRZ tests (`expected_results[i][j] = cp_t{..}` tables
indexed by `index (+ angles.size())`, applied to |0..0>)
plus the symbolic tables of the fixed-gate tests (Test_StateVectorKokkos_NonParam.cpp: SWAP, CZ, Toffoli,
CSWAP on H(0) X(1)|000>, rows written with z = ZERO and i = INVSQRT2; PauliY, PauliZ, S, T on |+++>, rows of
symbols defined from Util:: constants -- this script evaluates the symbols)
into tests/golden/ref_param_literals.json. Runs only where /root/reference exists (the build
container); the JSON it writes is committed and travels to the GPU box.

usage: python tests/golden/make_ref_literals.py
"""
import json
import os
import re

SRC = "/root/reference/pennylane_lightning_kokkos/src/tests/Test_StateVectorKokkos_Param.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_param_literals.json")

CP = re.compile(r"cp_t\{\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\}")
APPLY = re.compile(r"kokkos_sv\.apply([A-Z][A-Za-z]*)\(\s*\{([0-9,\s]*)\}\s*,\s*(true|false)\s*,\s*\{([-+0-9.eE,\s]*)\}\s*\)")


def block(text, start):
    """The brace-balanced initialiser that starts at text[start] == '{'."""
    depth, i = 0, start
    while True:
        depth += text[i] == "{"
        depth -= text[i] == "}"
        i += 1
        if depth == 0:
            return text[start:i], i


ASSIGN = re.compile(r"(expected_results(?:_adj)?)\[(\d+)\]\[(\d+)\]\s*=\s*cp_t\{\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\}")
APPLY_IDX = re.compile(r"kokkos_sv\.apply([A-Z][A-Za-z]*)\(\s*\{([0-9,\s]*)\}\s*,\s*(true|false)\s*,\s*\{\s*angles\[index\]\s*\}\s*\)")
USE = re.compile(r"(?:real|imag)\(\s*(expected_results(?:_adj)?)\[index(\s*\+\s*angles\.size\(\))?\]\[j\]")


def table_cases(text):
    """Test cases that start from |0..0>, loop `index` over `angles` and compare with a row of a
    sparse literal table."""
    out = []
    starts = [m.start() for m in re.finditer(r"TEMPLATE_TEST_CASE\(", text)] + [len(text)]
    for s0, s1 in zip(starts, starts[1:]):
        body = text[s0:s1]
        if "expected_results[" not in body or "ini_st" in body:
            continue
        nqm = re.search(r"num_qubits\s*=\s*(\d+)", body)
        ang = re.search(r"angles\s*=?\s*\{([-+0-9.eE,\s]*)\}", body)
        if not nqm or not ang:
            continue
        nq = int(nqm.group(1))
        angles = [float(a) for a in ang.group(1).split(",") if a.strip()]
        tables = {}
        for name, i, j, re_, im_ in ASSIGN.findall(body):
            tables.setdefault(name, {}).setdefault(int(i), {})[int(j)] = [float(re_), float(im_)]
        if not tables:
            # dense literal rows: expected_results{ std::vector<cp_t>{cp_t{..}, ..}, .. }
            for d in re.finditer(r"(expected_results(?:_adj)?)\s*\{", body):
                init, _ = block(body, d.end() - 1)
                rows = re.findall(r"std::vector<cp_t>\s*\{((?:\s*cp_t\{[^{}]*\}\s*,?)+)\}", init)
                if rows and all(len(CP.findall(r)) == 1 << nq for r in rows):
                    tables[d.group(1)] = {i: {j: [float(x), float(y)] for j, (x, y) in enumerate(CP.findall(r))}
                                          for i, r in enumerate(rows)}
        if not tables:
            continue
        for a in APPLY_IDX.finditer(body):
            u = USE.search(body, a.end())
            if not u or u.start() - a.end() > 500:
                continue
            if u.group(1) not in tables:
                continue
            for index, angle in enumerate(angles):
                row = index + (len(angles) if u.group(2) else 0)
                exp = [[0.0, 0.0] for _ in range(1 << nq)]
                for j, v in tables[u.group(1)][row].items():
                    exp[j] = v
                ini = [[0.0, 0.0] for _ in range(1 << nq)]
                ini[0] = [1.0, 0.0]
                out.append({"gate": a.group(1), "wires": [int(w) for w in a.group(2).split(",") if w.strip()],
                            "inverse": a.group(3) == "true", "params": [angle], "ini": ini, "expected": exp,
                            "ref_line": text.count("\n", 0, s0 + a.start()) + 1})
    return out


SRC_NP = "/root/reference/pennylane_lightning_kokkos/src/tests/Test_StateVectorKokkos_NonParam.cpp"
ROW_ZI = re.compile(r"const\s+std::vector<cp_t>\s+expected_results\s*=\s*\{([-zi,\s]*)\}\s*;")
APPLY_NP = re.compile(r"\.apply(?:([A-Z][A-Za-z]*)\(|Operation\(\s*\"([A-Za-z]+)\"\s*,)\s*\{([0-9,\s]*)\}\s*,\s*(true|false)\s*\)")
PREP = re.compile(r'applyOperation\(\s*\{\{"Hadamard"\},\s*\{"PauliX"\}\},\s*\{\{0\},\s*\{1\}\},\s*\{\{false\},\s*\{false\}\}\)')


def nonparam_cases(text):
    """Fixed gates on |+10> (H on wire 0, X on wire 1: amplitude 1/sqrt2 at indices 2 and 6, the test's
    own "Check Initial value" section) against rows of z / i symbols."""
    out, seen = [], set()
    h = 2.0 ** -0.5
    starts = [m.start() for m in re.finditer(r"TEMPLATE_TEST_CASE\(", text)] + [len(text)]
    for s0, s1 in zip(starts, starts[1:]):
        body = text[s0:s1]
        nqm = re.search(r"num_qubits\s*=\s*(\d+)", body)
        if not PREP.search(body) or not nqm or int(nqm.group(1)) != 3:
            continue
        ini = [[0.0, 0.0] for _ in range(8)]
        ini[2] = ini[6] = [h, 0.0]
        rows = list(ROW_ZI.finditer(body))
        for r, nxt in zip(rows, [x.start() for x in rows[1:]] + [len(body)]):
            sym = [t.strip() for t in r.group(1).split(",") if t.strip()]
            if len(sym) != 8:
                continue
            exp = [{"i": [h, 0.0], "-i": [-h, 0.0], "z": [0.0, 0.0]}[t] for t in sym]
            for a in APPLY_NP.finditer(body, r.end(), nxt):
                gate = a.group(1) or a.group(2)
                wires = [int(w) for w in a.group(3).split(",") if w.strip()]
                key = (gate, tuple(wires), a.group(4), tuple(sym))
                if gate == "Operation" or key in seen:
                    continue
                seen.add(key)
                out.append({"gate": gate, "wires": wires, "inverse": a.group(4) == "true", "params": [],
                            "ini": ini, "expected": exp, "ref_file": "Test_StateVectorKokkos_NonParam.cpp",
                            "ref_line": text.count("\n", 0, s0 + a.start()) + 1})
    return out


UTIL = {"HALF": 0.5, "INVSQRT2": 2.0 ** -0.5, "IMAG": 1j, "NEGONE": -1.0, "ONE": 1.0, "ZERO": 0.0}
SYMDEF = re.compile(r"(?:const\s+)?auto\s+([a-z])\s*=\s*([^;]*Util::[^;]*);")
PREP3H = re.compile(r'\{\{"Hadamard"\},\s*\{"Hadamard"\},\s*\{"Hadamard"\}\},\s*\{\{0\},\s*\{1\},\s*\{2\}\}')


def plus_state_cases(text):
    """PauliY / PauliZ / S / T on |+++>: rows of one-letter symbols defined from Util:: constants
    (`auto p = Util::HALF<..>() * Util::INVSQRT2<..>() * ...`), row `index` = gate on wire `index`."""
    out = []
    starts = [m.start() for m in re.finditer(r"TEMPLATE_TEST_CASE\(", text)] + [len(text)]
    for s0, s1 in zip(starts, starts[1:]):
        body = text[s0:s1]
        tab = re.search(r"std::vector<std::vector<cp_t>>\s+expected_results\s*=\s*\{", body)
        app = re.search(r"kokkos_sv\.apply([A-Z][A-Za-z]*)\(\s*\{index\}\s*,\s*(true|false)\s*\)", body)
        if not PREP3H.search(body) or not tab or not app:
            continue
        env = {}
        for name, expr in SYMDEF.findall(body[:tab.start()]):
            expr = re.sub(r"Util::([A-Z0-9]+)<[^>]*>\(\)", lambda m: repr(UTIL[m.group(1)]), expr)
            env[name] = complex(eval(" ".join(expr.split()), {"__builtins__": {}}, dict(env)))
        init, _ = block(body, tab.end() - 1)
        rows = re.findall(r"\{([a-z,\s]+)\}", init)
        if len(rows) != 3:
            continue
        amp = 0.5 * 2.0 ** -0.5
        for w, row in enumerate(rows):
            sym = [t.strip() for t in row.split(",") if t.strip()]
            if len(sym) != 8:
                break
            out.append({"gate": app.group(1), "wires": [w], "inverse": app.group(2) == "true", "params": [],
                        "ini": [[amp, 0.0]] * 8, "expected": [[env[t].real, env[t].imag] for t in sym],
                        "ref_file": "Test_StateVectorKokkos_NonParam.cpp",
                        "ref_line": text.count("\n", 0, s0 + app.start()) + 1})
    return out


def main():
    text = open(SRC).read()
    cases = []
    for m in re.finditer(r"std::vector<cp_t>\s+ini_st\s*\{", text):
        ini_txt, end = block(text, m.end() - 1)
        e = re.compile(r"std::vector<cp_t>\s+expected\s*\{").search(text, end)
        if not e or e.start() - end > 200:
            continue
        exp_txt, end2 = block(text, e.end() - 1)
        a = APPLY.search(text, end2)
        if not a or a.start() - end2 > 600:
            continue
        ini = [[float(x), float(y)] for x, y in CP.findall(ini_txt)]
        exp = [[float(x), float(y)] for x, y in CP.findall(exp_txt)]
        if len(ini) != len(exp) or len(ini) & (len(ini) - 1):
            continue
        line = text.count("\n", 0, m.start()) + 1
        cases.append({"gate": a.group(1), "wires": [int(w) for w in a.group(2).split(",") if w.strip()],
                      "inverse": a.group(3) == "true",
                      "params": [float(p) for p in a.group(4).split(",") if p.strip()],
                      "ini": ini, "expected": exp, "ref_line": line})
    cases += table_cases(text)
    cases += nonparam_cases(open(SRC_NP).read())
    cases += plus_state_cases(open(SRC_NP).read())
    with open(OUT, "w") as f:
        json.dump({"source": "reference src/tests/Test_StateVectorKokkos_Param.cpp, Test_StateVectorKokkos_NonParam.cpp", "cases": cases}, f, indent=0)
    print(f"{len(cases)} cases ->", OUT)
    for c in cases:
        print(" ", c["gate"], c["wires"], c["inverse"], c["params"], "line", c["ref_line"], len(c["ini"]))


if __name__ == "__main__":
    main()
