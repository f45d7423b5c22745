#!/usr/bin/env python
"""ncu metric log of ONE C4 loop pass (tools/r2_final.sh: gpu__time_duration.sum and the executed fp64 thread instructions
smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum of every kernel) -> profiles/r2_c4_fp64_ops.json, the op
count tools/bench_configs.py divides by the measured pass time for C4's fp64 roofline.

    python tools/c4_ops.py gpurun_out/<run>/c4_pass_counters.csv MIXED_CELLS [OUT.json]
"""
import collections
import csv
import json
import os
import re
import sys


def main():
    path, mixed = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_c4_fp64_ops.json")
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    kernels = collections.OrderedDict()
    per_id = collections.defaultdict(dict)
    for r in rows:
        per_id[r["ID"]]["name"] = r["Kernel Name"]
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        per_id[r["ID"]][r["Metric Name"]] = v
    total_ops = total_us = 0.0
    for k, d in per_id.items():
        name = re.sub(r"\(.*", "", d["name"]).strip()
        name = name.replace("vofx::", "")
        e = kernels.setdefault(name, {"launches": 0, "time_us": 0.0, "fp64_thread_inst": 0.0})
        ops = sum(d.get(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum", 0.0) for o in ("dfma", "dmul", "dadd"))
        e["launches"] += 1
        e["time_us"] = round(e["time_us"] + d.get("gpu__time_duration.sum", 0.0), 1)
        e["fp64_thread_inst"] += ops
        total_ops += ops
        total_us += d.get("gpu__time_duration.sum", 0.0)
    res = {"workload": "c4: 3D LeVeque sphere 256^3, iterative Swartz (maxIterations 10), pass 4", "mixed_cells": mixed,
           "fp64_thread_instructions_per_pass": total_ops, "serialised_kernel_time_us": total_us, "kernels": kernels,
           "source": "ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum, gpu__time_duration.sum --clock-control none (tools/r2_final.sh, tools/c4_ops.py)"}
    json.dump(res, open(out, "w"), indent=1)
    print(out, total_ops, total_us)


if __name__ == "__main__":
    main()
