#!/usr/bin/env python
"""Executed warp instructions and stall samples of one ncu report (source page), grouped by SASS opcode class.

    python tools/ncu_opcodes.py REPORT.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = rows[1]
    iS, iE, iN, iT = head.index("Source"), head.index("Instructions Executed"), head.index("# Samples"), head.index("Thread Instructions Executed")
    ex, sm, th = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= iT:
            continue
        toks = r[iS].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0].split(".")[0] if toks else "?"
        ex[op] += int(r[iE] or 0)
        sm[op] += int(r[iN] or 0)
        th[op] += int(r[iT] or 0)
    tot, tots = sum(ex.values()), sum(sm.values())
    print(f"{'opcode':12s} {'executed':>12s} {'%':>6s} {'samples':>8s} {'%':>6s} {'thr/inst':>8s}")
    for op, n in ex.most_common(40):
        print(f"{op:12s} {n:12d} {100*n/tot:6.2f} {sm[op]:8d} {100*sm[op]/max(tots,1):6.2f} {th[op]/max(n,1):8.1f}")
    print("total", tot, tots)


if __name__ == "__main__":
    main()
