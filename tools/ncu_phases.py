#!/usr/bin/env python
"""Executed warp instructions and stall samples of a kernel per source line of ONE file, with inlined callees folded
into the line of that file they were inlined at (nvdisasm -gi), optionally summed over named line ranges.

    python tools/ncu_phases.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING FILE [name:first-last ...]

e.g. the phases of the plane-constant solve: FILE = vofx_coop3d.cuh, ranges rotate:702-725 cellvol:726-767 ..."""
import collections
import csv
import io
import re
import subprocess
import sys


def frames_per_address(cubin, kernel):
    text = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    out, inside, chain = {}, False, []
    pending = []
    for line in text.splitlines():
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                chain = pending
            pending = []
            out[int(m.group(1), 16)] = list(chain)
        elif line.startswith(".section") or line.startswith("//----"):
            inside = False
    return out


def main():
    rep, cubin, kernel, fname = sys.argv[1:5]
    ranges = []
    for a in sys.argv[5:]:
        name, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((name, int(lo), int(hi)))
    frames = frames_per_address(cubin, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = rows[1]
    iA, iE, iN, iS = head.index("Address"), head.index("Instructions Executed"), head.index("# Samples"), head.index("Source")
    iT = head.index("Thread Instructions Executed")
    base = int(rows[2][iA], 16)
    ex, sm = collections.Counter(), collections.Counter()
    stall_cols = [(i, h) for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.defaultdict(collections.Counter)
    fp = collections.Counter()
    th = collections.Counter()
    for r in rows[2:]:
        if len(r) <= iN:
            continue
        off = int(r[iA], 16) - base
        chain = frames.get(off, [])
        key = ("?", 0)
        for f, l in chain:                     # innermost first; keep the OUTERMOST frame inside the chosen file
            if f == fname:
                key = (f, l)
        if key[0] == "?" and chain:
            key = chain[-1]
        n = int(r[iE] or 0)
        ex[key] += n
        th[key] += int(r[iT] or 0)
        for i, h in stall_cols:
            if i < len(r) and r[i]:
                stalls[key][h[6:]] += int(r[i] or 0)
        sm[key] += int(r[iN] or 0)
        toks = r[iS].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if toks and toks[0].split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"):
            fp[key] += n
    tot, tots = sum(ex.values()), sum(sm.values())
    if ranges:
        print(f"{'phase':16s} {'executed':>12s} {'%':>6s} {'fp64 %':>7s} {'samples':>8s} {'%':>6s} {'thr/inst':>8s}")
        rest_e = rest_s = 0
        used = set()
        for name, lo, hi in ranges:
            keys = [k for k in ex if k[0] == fname and lo <= k[1] <= hi]
            used.update(keys)
            e, s, f, t = sum(ex[k] for k in keys), sum(sm[k] for k in keys), sum(fp[k] for k in keys), sum(th[k] for k in keys)
            st = collections.Counter()
            for k in keys:
                st.update(stalls[k])
            top = " ".join(f"{n}:{100*v/max(s,1):.0f}%" for n, v in st.most_common(3))
            print(f"{name:16s} {e:12d} {100*e/tot:6.2f} {100*f/max(e,1):7.1f} {s:8d} {100*s/max(tots,1):6.2f} {t/max(e,1):8.1f}  {top}")
        others = collections.Counter()
        for k in ex:
            if k not in used:
                others[k] += ex[k]
        e = sum(others.values())
        print(f"{'(other)':16s} {e:12d} {100*e/tot:6.2f}")
        for k, n in others.most_common(12):
            print(f"    {k[0]}:{k[1]}  {n}  {100*n/tot:.2f}%")
    else:
        for k, n in sorted(ex.items(), key=lambda kv: -kv[1])[:60]:
            print(f"{k[0]}:{k[1]:<6d} {n:12d} {100*n/tot:6.2f} {sm[k]:8d} {100*sm[k]/max(tots,1):6.2f}")
    print("total", tot, tots)


if __name__ == "__main__":
    main()
