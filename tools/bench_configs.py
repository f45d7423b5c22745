#!/usr/bin/env python
"""Secondary BASELINE.json configs, measured like bench.py measures configs[2] (one JSON line per config):

    c1  2D rotating circle 128^2, Young's                       (configs[0], plumbing)
    c2  2D Zalesak (SlottedCylinder) 4096^2, HF + curvature      (configs[1])
    c4  3D sphere in the LeVeque field 256^3, iterative Swartz  (configs[3])

bench.py stays the headline (configs[2]); these lines go to profiles/.  CUDA events, inputs resident in HBM,
per-kernel times from the C ABI's profiling hooks, the reference's CPU templates on a bounded sample beside it."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_sample(dim, n, recon, problem, passes, warm):
    from oracle import oraclelib as O
    from oracle import reflib as R
    mod, grid = (R, R.RefGrid((n,) * dim)) if R.available() else (O, O.OracleGrid((n,) * dim))
    u = grid.init(problem)
    t = dt = 0.0
    u, t, dt, _, _ = grid.run(recon, problem, 1e-6, 0.5, warm, u, t, dt)
    t0 = time.perf_counter()
    grid.run(recon, problem, 1e-6, 0.5, passes, u, t, dt)
    wall = time.perf_counter() - t0
    return {"value": n ** dim * passes / wall, "unit": "cell-steps/s", "cores": 1, "kind": "reference" if mod is R else "port",
            "sample": f"{n}^{dim}, {passes} passes after {warm}", "seconds": wall}


def run(name, n, problem, recon, curvature, steps, warmup, cpu):
    import torch
    import dune_vof_b200.operators as ops
    from dune_vof_b200 import lib as vlib
    grid = ops.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    grid.use_current_stream()
    if len(n) == 2 and problem == ops.PROB_ROTATING_CIRCLE:
        uh = grid.empty("color")                                # exact circle average (test/average.hh:82-101) on the device
        vlib.check(grid._L.vofx_init_circle(grid._ctx, (C.c_double * 2)(0.5, 0.75), 0.15, grid.pc(uh)))
    else:
        uh = grid.init(problem)
    alg = ops.Algorithm(grid, 0.5, 1e-6, recon, with_curvature=curvature, incremental=True)
    alg.begin(0.0, 0.0)
    graph = True
    from bench import ClockSampler, measured_peaks
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        grid.use_current_stream()
        for _ in range(warmup):
            alg.step(uh)
    torch.cuda.synchronize()
    if graph:
        alg.capture(uh)
        one_step = alg.replay
    else:
        grid.use_current_stream()
        one_step = lambda: alg.step(uh)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    grid.use_current_stream()
    L = grid._L
    vlib.check(L.vofx_profile_enable(grid._ctx, 1))
    kp = min(steps, 20)
    for _ in range(kp):
        alg.step(uh)
    names = C.create_string_buffer(48 * 32)
    tot = (C.c_double * 32)()
    cnt = (C.c_int64 * 32)()
    nent = C.c_int(0)
    vlib.check(L.vofx_profile_read(grid._ctx, 32, names, tot, cnt, C.byref(nent)))
    vlib.check(L.vofx_profile_enable(grid._ctx, 0))
    kernels = {names.raw[48 * i:48 * (i + 1)].split(b"\0", 1)[0].decode(): round(tot[i] / max(cnt[i], 1), 5) for i in range(nent.value)}
    sampler.stop()
    st = alg.state()
    cells = 1
    for x in n:
        cells *= x
    bytes_per_cell = 25 if curvature else 17
    peak, peak_src = measured_peaks()
    value = cells * steps / (ms * 1e-3)
    if len(n) == 3 and recon == ops.RECON_SWARTZ:
        # fp64-bound: executed fp64 thread instructions per pass (static: the ncu counters of one pass of this
        # workload, profiles/r2_c4_fp64_ops.json) over the measured pass time, against the measured fp64 issue rate of this GPU
        rates = (C.c_double * 2)()
        vlib.check(grid._L.vofx_test_fp64_rate(rates))
        roofline = {"bound": "fp64", "peak": rates[1] / 1e3, "unit": "T fp64 instructions/s (separate multiply / add, as the parity contract compiles)",
                    "peak_fma": rates[0] / 1e3, "peak_source": "vofx_test_fp64_rate, measured in this run", "achieved": None, "frac": None, "traffic": None}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r2_c4_fp64_ops.json")))
            per_mixed = prof["fp64_thread_instructions_per_pass"] / prof["mixed_cells"]
            achieved = per_mixed * st.mixed_cells / (ms / steps * 1e-3) / 1e12
            roofline.update({"achieved": achieved, "frac": achieved / (rates[1] / 1e3),
                             "fp64_thread_instructions_per_mixed_cell": per_mixed,
                             "ops_source": "profiles/r2_c4_fp64_ops.json (ncu smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on of one pass)"})
        except Exception as exc:
            roofline["note"] = "no op counts: " + repr(exc)
    else:
        roofline = {"bound": "hbm", "achieved": value * bytes_per_cell / 1e9, "peak": peak, "unit": "GB/s", "frac": value * bytes_per_cell / 1e9 / peak,
                    "algorithmic_bytes_per_cell_step": bytes_per_cell, "peak_source": peak_src, "traffic": None}
    line = {"config": name, "grid": list(n), "reconstruction": int(recon), "curvature": bool(curvature), "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "launch": "CUDA graph replay" if graph else "eager", "value": value, "unit": "cell-steps/s",
            "flagging": "incremental", "clocks": sampler.report(), "roofline": roofline,
            "mixed_cells": st.mixed_cells, "kernel_avg_ms": kernels, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if hasattr(alg, "_graph"):
        del alg._graph
    torch.cuda.synchronize()
    grid.close()
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c4")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import dune_vof_b200.operators as ops
    from oracle import oraclelib as O
    for c in a.configs.split(","):
        if c == "c1":
            cpu = None if a.no_cpu else cpu_sample(2, 128, O.RECON_YOUNGS, O.PROB_ROTATING_CIRCLE, 100, 5)
            run("c1: 2D rotating circle 128^2, Young's", (128, 128), ops.PROB_ROTATING_CIRCLE, ops.RECON_YOUNGS, False, 200, 10, cpu)
        elif c == "c2":
            cpu = None if a.no_cpu else cpu_sample(2, 1024, O.RECON_HF, O.PROB_SLOTTED_CYLINDER, 20, 2)
            run("c2: 2D Zalesak 4096^2, HF + curvature", (4096, 4096), ops.PROB_SLOTTED_CYLINDER, ops.RECON_HF, True, 200, 10, cpu)
        elif c == "c3hf":
            run("c3hf: 3D LeVeque sphere 512^3, height functions (the reference's default factory)", (512, 512, 512), ops.PROB_LEVEQUE, ops.RECON_HF, False, 50, 5, None)
        elif c == "c4":
            cpu = None if a.no_cpu else cpu_sample(3, 32, O.RECON_SWARTZ, O.PROB_LEVEQUE, 2, 1)
            run("c4: 3D LeVeque sphere 256^3, iterative Swartz", (256, 256, 256), ops.PROB_LEVEQUE, ops.RECON_SWARTZ, False, 20, 3, cpu)


if __name__ == "__main__":
    main()
    sys.stdout.flush()
