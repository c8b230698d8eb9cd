#!/usr/bin/env python
"""Extracts the known-answer measurement tests of the reference (circuit from |0..0>, one observable, one
expected number) into tests/golden/ref_measure_kats.json:

  src/tests/Test_StateVectorKokkos_Expval.cpp:19-336  expval of Identity / PauliX / PauliY / PauliZ / Hadamard
                                                      (direct functor call and NamedObs), 1- and 2-qubit matrices
                                                      (direct call and HermitianObs)
  src/tests/Test_StateVectorKokkos_Var.cpp:19-122     var of NamedObs, HermitianObs, TensorProdObs

Runs only where /root/reference exists (the build container); the JSON travels to the GPU box.
usage: python tests/golden/make_ref_measure_kats.py
"""
import json
import math
import os
import re

TESTS = "/root/reference/pennylane_lightning_kokkos/src/tests/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_measure_kats.json")
ENV = {"PI": math.pi, "M_PI": math.pi}

OP1 = re.compile(r'kokkos_sv\.applyOperation\(\s*"(\w+)"\s*,\s*\{([0-9,\s]*)\}\s*,\s*(true|false)\s*(?:,\s*\{([^}]*)\})?\s*\)')
OP2 = re.compile(r'kokkos_sv\.apply([A-Z]\w*)\(\s*\{([0-9,\s]*)\}\s*(?:,\s*(true|false))?\s*(?:,\s*\{([^}]*)\})?\s*\)')
DIRECT = re.compile(r'getExpectationValue(Identity|PauliX|PauliY|PauliZ|Hadamard)\(\s*\{([0-9,\s]*)\}\s*\)')
NAMED = re.compile(r'NamedObsKokkos<TestType>>?\(\s*"(\w+)"\s*,\s*(?:std::vector<size_t>)?\{([0-9,\s]*)\}\s*\)')
MATW = re.compile(r'(?:HermitianObsKokkos<TestType>\(\s*matrix|getExpectationValue(?:Single|Two)QubitOp\(\s*opMatDevice)\s*,\s*\{([0-9,\s]*)\}\s*\)')
ASSIGN = re.compile(r'(?:matrix|opMat)\[(\d+)\]\s*=\s*([^;]+);')
INITL = re.compile(r'matrix\{(.*?)\};', re.S)


def ints(txt):
    return [int(t) for t in txt.split(",") if t.strip()]


def cplx(expr, env):
    e = re.sub(r"Kokkos::complex(?:<TestType>)?|static_cast<TestType>", "", expr).strip()
    e = e.replace("{", "(").replace("}", ")")
    v = eval(e, {"__builtins__": {}}, env)
    while isinstance(v, tuple) and len(v) == 1:
        v = v[0]
    if isinstance(v, tuple):
        a, b = v
        a = a[0] if isinstance(a, tuple) else a
        return [float(a), float(b)]
    return [float(v), 0.0]


def sections(body):
    """(offset, text) of the innermost SECTION blocks."""
    out = []
    for m in re.finditer(r"SECTION\(", body):
        i = body.index("{", m.end())
        depth, j = 0, i
        while True:
            depth += body[j] == "{"
            depth -= body[j] == "}"
            j += 1
            if depth == 0:
                break
        txt = body[i:j]
        if "SECTION(" not in txt[1:]:
            out.append((m.start(), txt))
    return out


def extract(fname, lo, hi):
    text = open(TESTS + fname).read()
    lines = text.split("\n")
    start = sum(len(l) + 1 for l in lines[:lo - 1])
    end = sum(len(l) + 1 for l in lines[:hi])
    cases = []
    starts = [m.start() for m in re.finditer(r"TEMPLATE_TEST_CASE\(", text)] + [len(text)]
    for s0, s1 in zip(starts, starts[1:]):
        if s0 < start or s0 >= end:
            continue
        body = text[s0:s1]
        n = int(re.search(r"num_qubits\s*=\s*(\d+)", body).group(1))
        head = body[:body.index("SECTION(")]
        for off, sec in sections(body):
            src = head + sec
            ops = []
            for m in sorted(list(OP1.finditer(src)) + list(OP2.finditer(src)), key=lambda m: m.start()):
                name, wires, inv, par = m.group(1), ints(m.group(2)), m.group(3) == "true", m.group(4)
                if name == "Operation":
                    continue
                params = [float(eval(p, {"__builtins__": {}}, ENV)) for p in par.split(",")] if par and par.strip() else []
                ops.append([name, wires, inv, params])
            env = dict(ENV)
            if "theta" in sec:
                env["theta"] = math.pi / 2
                env["c"] = math.cos(env["theta"] / 2)
                env["js"] = math.sin(-env["theta"] / 2)
            obs = None
            mw = MATW.search(sec)
            named = NAMED.findall(sec)
            if mw:
                wires = ints(mw.group(1))
                dim = 1 << len(wires)
                mat = [[0.0, 0.0] for _ in range(dim * dim)]
                il = INITL.search(sec)
                if il:
                    parts = re.findall(r"\{[^{}]*\}|[^,{}\s]+", il.group(1))
                    mat = [cplx(p, env) for p in parts]
                for idx, expr in ASSIGN.findall(sec):
                    mat[int(idx)] = cplx(expr, env)
                assert len(mat) == dim * dim
                obs = {"type": "hermitian", "matrix": mat, "wires": wires}
            elif DIRECT.search(sec):
                d = DIRECT.search(sec)
                obs = {"type": "named", "name": d.group(1), "wires": ints(d.group(2)), "call": "direct"}
            elif len(named) == 1:
                obs = {"type": "named", "name": named[0][0], "wires": ints(named[0][1]), "call": "obs"}
            elif len(named) > 1 and "TensorProdObsKokkos" in sec:
                obs = {"type": "tensor", "factors": [{"name": a, "wires": ints(b)} for a, b in named]}
            if obs is None:
                continue
            kind = "var" if re.search(r"\bm\.var\(", sec) else "expval"
            chk = re.search(r"CHECK\(([^;]*)\);", sec).group(1)
            em = re.search(r"expected\s*=\s*TestType\(([-+0-9.eE]+)\)", sec)
            if em:
                expected, exact = float(em.group(1)), False
            else:
                rhs = chk.replace("res", "").replace("==", "").strip()
                neg = rhs.startswith("-")
                sym = re.sub(r"[-()\s]|Approx", "", rhs)
                val = {"ONE": 1.0, "ZERO": 0.0, "0": 0.0, "INVSQRT2": 0.707106781186547524401}[sym]
                expected, exact = (-val if neg else val), "Approx" not in rhs
            cases.append({"kind": kind, "n": n, "ops": ops, "obs": obs, "expected": expected,
                          "exact_compare": exact, "ref_file": fname,
                          "ref_line": text.count("\n", 0, s0 + off) + 1})
    return cases


def main():
    cases = extract("Test_StateVectorKokkos_Expval.cpp", 19, 336) + extract("Test_StateVectorKokkos_Var.cpp", 19, 122)
    with open(OUT, "w") as f:
        json.dump({"source": "reference src/tests/Test_StateVectorKokkos_Expval.cpp:19-336, Test_StateVectorKokkos_Var.cpp:19-122",
                   "cases": cases}, f, indent=0)
    print(len(cases), "cases ->", OUT)
    for c in cases:
        o = c["obs"]
        print(" ", c["ref_file"][22:-4], c["ref_line"], c["kind"], o["type"], o.get("name", ""), o.get("wires", ""),
              len(c["ops"]), "ops ->", c["expected"])


if __name__ == "__main__":
    main()
