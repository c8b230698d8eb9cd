#!/usr/bin/env python
"""Extracts the literal input / expected state vectors of the reference's own parametric-gate tests
(reference src/tests/Test_StateVectorKokkos_Param.cpp: blocks `std::vector<cp_t> ini_st{...}` +
`std::vector<cp_t> expected{...}` followed by `kokkos_sv.apply<Gate>({wires}, inverse, {params})`)
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
    with open(OUT, "w") as f:
        json.dump({"source": "reference src/tests/Test_StateVectorKokkos_Param.cpp", "cases": cases}, f, indent=0)
    print(f"{len(cases)} cases ->", OUT)
    for c in cases:
        print(" ", c["gate"], c["wires"], c["inverse"], c["params"], "line", c["ref_line"], len(c["ini"]))


if __name__ == "__main__":
    main()
