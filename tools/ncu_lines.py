#!/usr/bin/env python
"""Join the per-instruction page of an ncu report with the line table of the cubin (nvdisasm -g) and print where the
warp-stall samples and the executed instructions of a kernel fall, per source line.

    python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [TOP]

(ncu's own CSV export of the source page has no line column.)  Used for the band kernels' hot-spot lists in profiles/."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, inside, where = None, False, ("?", 0)
        for line in text.splitlines():
            if line.startswith(".text."):
                inside = kernel in line
                if inside:
                    cur = line
                    out = []
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                where = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), where, m.group(2).strip()))
            elif line.startswith(".section") or line.startswith("//----"):
                inside = False
        if out:
            return out
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    si, ei, so = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    inst = [(int(r[si] or 0), int(r[ei] or 0), r[so].strip()) for r in rows[hdr + 1:] if len(r) > si]
    sass = sass_lines(lib, kernel)
    if len(sass) != len(inst):
        print(f"warning: {len(sass)} instructions in the cubin, {len(inst)} in the report (different builds?)", file=sys.stderr)
    n = min(len(sass), len(inst))
    by_line = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_e = 0
    for k in range(n):
        s, e, _ = inst[k]
        _, where, _ = sass[k]
        by_line[where][0] += s
        by_line[where][1] += e
        tot_s += s
        tot_e += e
    print(f"{kernel}: {n} SASS instructions, {tot_s} samples, {tot_e} warp instructions executed")
    print(f"{'file:line':34s} {'samples':>8s} {'%':>6s} {'executed':>10s} {'%':>6s}")
    for where, (s, e) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{where[0] + ':' + str(where[1]):34s} {s:8d} {100 * s / max(tot_s, 1):6.2f} {e:10d} {100 * e / max(tot_e, 1):6.2f}")


if __name__ == "__main__":
    main()
