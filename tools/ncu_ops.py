#!/usr/bin/env python
"""opcode histogram (stall samples, executed warp instructions) of the first kernel of an ncu report"""
import collections, csv, re, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ei, so = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
samp, exe = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= si:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[so].strip())
    op = (m.group(2) if m else r[so]).split(".")[0]
    samp[op] += int(r[si] or 0)
    exe[op] += int(r[ei] or 0)
ts, te = sum(samp.values()), sum(exe.values())
print(f"samples {ts}, warp instructions {te}")
for op, c in exe.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} executed {c:10d} {100 * c / te:5.1f} %   samples {samp[op]:6d} {100 * samp[op] / ts:5.1f} %")
