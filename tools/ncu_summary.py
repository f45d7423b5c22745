"""Summarise `ncu --set full` reports (one kernel launch each) into profiles/<tag>_ncu_summary.json and
profiles/<tag>_ncu_<kernel>_details.txt.

    python tools/ncu_summary.py gpurun_out/r1g_ r1 [OUTDIR]

reads gpurun_out/r1g_<kernel>.ncu-rep for every kernel found and writes into OUTDIR (default profiles/; on the GPU box the
directory under gpurun_out/ that travels back, so that the reports themselves - tens of MB - need not).  An existing
summary in OUTDIR is extended.  The numbers bench.py's roofline.traffic cites come from here.
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    values = dict(zip(rows[0], rows[2]))
    UNITS.update(dict(zip(rows[0], rows[1])))
    return values


UNITS = {}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
         "byte/block": 1.0, "Kbyte/block": 1e3}


def num(d, k):
    """value in base units: microseconds for durations, bytes for sizes, the raw number otherwise"""
    try:
        return float(d[k].replace(",", "")) * SCALE.get(UNITS.get(k, ""), 1.0)
    except (KeyError, ValueError):
        return None


def main():
    prefix, tag = sys.argv[1], sys.argv[2]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outdir = sys.argv[3] if len(sys.argv) > 3 else os.path.join(root, "profiles")
    summary = {}
    if os.path.exists(os.path.join(outdir, f"{tag}_ncu_summary.json")):
        summary = json.load(open(os.path.join(outdir, f"{tag}_ncu_summary.json")))
    for rep in sorted(glob.glob(prefix + "*.ncu-rep")):
        kernel = os.path.basename(rep)[len(os.path.basename(prefix)):-len(".ncu-rep")]
        d = raw(rep)
        stalls = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(num(d, k), 2)
                  for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                  and num(d, k) is not None}
        top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
        inst = num(d, "smsp__inst_executed.sum")
        thr_per_inst = num(d, "smsp__thread_inst_executed_per_inst_executed.ratio")
        summary[kernel] = {
            "duration_us": num(d, "gpu__time_duration.sum"),
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": (rd or 0.0) + (wr or 0.0),
            "registers_per_thread": num(d, "launch__registers_per_thread"),
            "shared_bytes_per_block": num(d, "launch__shared_mem_per_block_static"),
            "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": inst,
            "ipc_active": num(d, "sm__inst_executed.avg.per_cycle_active"),
            "threads_per_warp_instruction": round(thr_per_inst, 2) if thr_per_inst else None,
            "fp64_pipe_active_pct": num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "stall_cycles_per_issue": top,
        }
        details = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        with open(os.path.join(outdir, f"{tag}_ncu_{kernel}_details.txt"), "w") as f:
            f.write(details)
    with open(os.path.join(outdir, f"{tag}_ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: {"us": v["duration_us"], "dram": v["dram_bytes_total"]} for k, v in summary.items()}))


if __name__ == "__main__":
    main()
