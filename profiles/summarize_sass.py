#!/usr/bin/env python
"""SASS summary of the built library (no GPU needed): per kernel, the instruction count and the mnemonics
that tell the data path (LDGSTS / UBLKCP loads, FFMA2 / DFMA arithmetic, spills, setmaxnreg ...).

usage: python profiles/summarize_sass.py [path/to/libb2sv.so] > profiles/<round>_tile_kernel.sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pennylane_lightning_kokkos_b200", "libb2sv.so")
KEYS = ["LDGSTS", "UBLKCP", "SYNCS", "DFMA", "DMUL", "FFMA2", "FFMA", "FMUL", "LDS", "STS", "STG", "LDG", "LDL", "STL",
        "USETMAXREG", "BAR", "NANOSLEEP", "LDC", "LDCU", "REDUX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    res = {}
    cur = None
    for line in usage.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            res[cur] = m.groups()
    print("# SASS summary (cuobjdump -sass / --dump-resource-usage of libb2sv.so, sm_100a)")
    print("# per kernel: registers / stack bytes / static shared bytes, instruction count, data-path mnemonics\n")
    name, ops = None, collections.Counter()

    def flush():
        if name and ("tile_exec" in name or "transition" in name or "csr_expval" in name or "exchange" in name):
            short = re.sub(r"^_ZN4b2sv|EEvPNS_4AmpTIT_E4typeENS_10PassParamsEmjj$", "", name)[:100]
            r = res.get(name, ("?", "?", "?"))
            print(short)
            print(f"  REG={r[0]} STACK={r[1]} SHARED={r[2]}  instructions {sum(ops.values())}  " +
                  "  ".join(f"{k}={ops[k]}" for k in KEYS if ops[k]))

    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, ops = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    flush()


if __name__ == "__main__":
    main()
