#!/usr/bin/env python
"""bench.py -- VoF cell-steps/sec of the per-timestep hot path (flag -> reconstruct -> flux/evolve -> update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl vofx|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json `metric`, configs[2]): 3D sphere (R = 0.15) in the LeVeque deformation field on a 512^3
Cartesian grid, Young's reconstruction, eps 1e-6, cfl 0.5, fp64.  A step is one pass of the time loop
(dune/vof/algorithm.hh:68-95 without l1error/output) over the whole grid.  N > 1 is WEAK scaling: every GPU owns one
512 x 512 x 512 slab of a 512 x 512 x 512N grid (z-slabs, 2-layer colour halos over NCCL, one MIN all-reduce); the
LeVeque field is 1-periodic in z and every slab starts with its own sphere, so all GPUs carry the same band load.

Prints ONE JSON line (rank 0).  `value` = cells x steps / device time (inputs resident in HBM), `e2e` = the same
through the host-buffer C-ABI call with H2D/D2H copies of the colour function inside the timed region,
`roofline` for the dominant kernel (dense flag/compact pass), `cpu_baseline` = the reference's own CPU templates
(oracle/_ref, or the plain-C port if that is absent) on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SIDE = int(os.environ.get("VOFX_BENCH_N", "512"))          # cells per axis of one slab
CPU_SAMPLE_N = int(os.environ.get("VOFX_BENCH_CPU_N", "128"))  # the reference needs >= 624 B/cell: 512^3 does not fit any host
EPS, CFL = 1e-6, 0.5
SURVEY_BYTES_PER_CELL_STEP = 17      # read u (8) + write u' (8) + write flag (1)
FLAG_KERNEL_BYTES_PER_CELL = 9       # what the dense kernel itself must move: read u (8) + write flag (1)
WORKLOAD_NAME = "3D sphere (R=0.15) in LeVeque deformation field, Young's reconstruction, eps 1e-6, cfl 0.5 (configs[2])"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """polls SM clock and throttle reasons through NVML while the timed region runs"""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2.0)

    def report(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_reference_run(passes, warm, sizes=(96, CPU_SAMPLE_N)):
    """The reference's own CPU implementation (oracle/_ref = /root/reference's templates, 1 thread: the reference has no
    threading and MPI is absent), or the plain-C port when the prebuilt harness is absent, on the bench problem.

    The reference stores >= 624 B/cell of stencil entities (24.7 GB at 256^3), so 512^3 does not fit a host.  It is
    therefore timed at TWO sample sizes of the same problem and every phase (flag / reconstruct / evolve / axpy) is
    fitted as  seconds = a * cells + b * mixed cells;  `value` is the fit evaluated on the bench grid (N_SIDE^3 cells and
    the mixed-cell count of that grid, scaled from the larger sample by the surface law (N_SIDE / n)^2), i.e. the
    reference on the bench's own config.  The raw samples are reported next to it.  Returns (value, info)."""
    from oracle import oraclelib as O
    from oracle import reflib as R
    import numpy as np
    if R.available():
        kind, Grid, mod = "reference", R.RefGrid, R
    else:
        kind, Grid, mod = "port", O.OracleGrid, O
    samples = []
    t_begin = time.perf_counter()
    for n_side in sizes:
        grid = Grid((n_side,) * 3)
        u = grid.init(mod.PROB_LEVEQUE)
        t = dt = 0.0
        if warm > 0:
            u, t, dt, _, _ = grid.run(mod.RECON_YOUNGS, mod.PROB_LEVEQUE, EPS, CFL, warm, u, t, dt)
        t0 = time.perf_counter()
        u, t, dt, phases, mixed = grid.run(mod.RECON_YOUNGS, mod.PROB_LEVEQUE, EPS, CFL, passes, u, t, dt)
        wall = time.perf_counter() - t0
        samples.append({"n": n_side, "cells": n_side ** 3, "passes": passes, "seconds": wall,
                        "cell_steps_per_s": n_side ** 3 * passes / wall, "mixed_cells_per_step": mixed / max(passes, 1),
                        "phase_seconds": {"flag": phases[0], "reconstruct": phases[1], "evolve": phases[2], "axpy": phases[3]}})
        grid.close()
        del grid, u
    lo, hi = samples[0], samples[-1]
    target_cells = N_SIDE ** 3
    target_mixed = hi["mixed_cells_per_step"] * (N_SIDE / hi["n"]) ** 2
    fit = {}
    step_seconds = 0.0
    if len(samples) >= 2 and lo["n"] != hi["n"]:
        A = np.array([[lo["cells"], lo["mixed_cells_per_step"]], [hi["cells"], hi["mixed_cells_per_step"]]], dtype=float)
        for name in ("flag", "reconstruct", "evolve", "axpy"):
            b = np.array([lo["phase_seconds"][name] / lo["passes"], hi["phase_seconds"][name] / hi["passes"]])
            per_cell, per_mixed = np.linalg.solve(A, b)
            per_cell, per_mixed = max(per_cell, 0.0), max(per_mixed, 0.0)
            fit[name] = {"ns_per_cell": per_cell * 1e9, "us_per_mixed_cell": per_mixed * 1e6}
            step_seconds += per_cell * target_cells + per_mixed * target_mixed
        other = [(smp["seconds"] - sum(smp["phase_seconds"].values())) / smp["passes"] / smp["cells"] for smp in samples]
        step_seconds += max(other[-1], 0.0) * target_cells
    else:
        step_seconds = hi["seconds"] / hi["passes"] * target_cells / hi["cells"]
    value = target_cells / step_seconds
    info = {"value": value, "unit": "cell-steps/s", "cores": 1, "kind": kind,
            "sample": f"{kind} timed at " + " and ".join(f"{smp['n']}^3" for smp in samples) + f" ({passes} loop passes after {warm} warm-up "
                      f"passes each, sphere in LeVeque field, Young's); per-phase fit seconds = a*cells + b*mixed evaluated at "
                      f"{N_SIDE}^3 cells / {target_mixed:.0f} mixed cells (the reference needs >= 624 B/cell: {N_SIDE}^3 does not fit a host)",
            "host_cores_available": os.cpu_count(),
            "threads_note": "the reference's path is single-threaded (no OpenMP / TBB; its parallelism is one MPI rank per core, and MPI is not in "
                            "this image): 1 core is what it can use here.  value x host cores would be the ceiling of an ideal MPI run",
            "ideal_all_cores_ceiling": value * (os.cpu_count() or 1),
            "seconds": time.perf_counter() - t_begin,
            "equivalent_seconds_per_step_on_bench_grid": step_seconds, "bench_grid_mixed_cells_assumed": target_mixed,
            "fit": fit, "samples": samples}
    return value, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 6)
    warm = 1
    sizes = (128, 256) if _mem_available_gb() > 48 else (96, CPU_SAMPLE_N)
    value, info = cpu_reference_run(steps, warm, sizes)
    # cross-check on the exact bench grid with the plain-C port (it needs 50 B/cell): measured, not fitted
    port = None
    if N_SIDE <= 512 and _mem_available_gb() > 24 and not args.no_port_check:
        try:
            from oracle import oraclelib as O
            og = O.OracleGrid((N_SIDE,) * 3)
            u = og.init(O.PROB_LEVEQUE)
            u, t, dt, _, _ = og.run(O.RECON_YOUNGS, O.PROB_LEVEQUE, EPS, CFL, 1, u)
            t0 = time.perf_counter()
            u, t, dt, ph, mixed = og.run(O.RECON_YOUNGS, O.PROB_LEVEQUE, EPS, CFL, 2, u, t, dt)
            wall = time.perf_counter() - t0
            port = {"kind": "port", "grid": [N_SIDE] * 3, "passes": 2, "seconds": wall, "cell_steps_per_s": N_SIDE ** 3 * 2 / wall,
                    "mixed_cells_per_step": mixed / 2, "cores": 1}
        except Exception as exc:
            port = {"error": repr(exc)}
    info["port_on_bench_grid"] = port
    line = {
        "impl": "reference",
        "metric": "VoF cell-steps/sec (reconstruct+advect)", "value": value, "unit": "cell-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * info["equivalent_seconds_per_step_on_bench_grid"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "grid": [N_SIDE] * 3,
                   "cpu_sample": info["sample"]},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vofx", choices=["vofx", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-port-check", action="store_true", help="reference arm: skip the plain-C port on the exact bench grid")
    ap.add_argument("--dense-flags", action="store_true", help="flag every cell every step instead of incremental flagging")
    ap.add_argument("--transport", default=None, choices=["peer", "nccl"], help="N > 1: peer-mapped windows (default) or NCCL halo exchange")
    ap.add_argument("--no-extra-workloads", action="store_true", help="N > 1: skip the c3-strong / c5 runs next to the weak-scaling line")
    ap.add_argument("--extra-steps", type=int, default=60, help="timed steps of the c3-strong / c5 runs")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout must carry exactly ONE JSON line: libraries that print there (NCCL's version banner at NCCL_DEBUG >= VERSION)
    # are sent to stderr for the duration of the run; the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(line, flush=True)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        emit(json.dumps({"error": "no CUDA device: vofx has no CPU fallback"}))
        return 2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION makes NCCL print its version banner on stdout, which must carry exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import dune_vof_b200.operators as ops
    from dune_vof_b200.decomposition import SlabAlgorithm, split_by_cost
    from dune_vof_b200 import lib as vlib
    L = vlib.load()

    W = max(3, args.warmup)
    K = max(1, args.steps)
    incremental = not args.dense_flags
    transport = args.transport if world > 1 else None
    use_graph = os.environ.get("VOFX_NO_GRAPH", "0") != "1"

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ints(x):
        if world == 1:
            return [int(x)]
        out = [None] * world
        dist.all_gather_object(out, int(x))
        return out

    def make_slab(n_global, h, layers=None, tr=transport, recon=ops.RECON_YOUNGS):
        slab = SlabAlgorithm(n_global, rank, world, CFL, EPS, recon, h=h, device=local_rank,
                             overlap=os.environ.get("VOFX_NO_OVERLAP", "0") != "1", incremental=incremental, transport=tr, layers=layers)
        slab.set_velocity_builtin(ops.PROB_LEVEQUE)
        slab.grid.use_current_stream()
        return slab

    def timed_run(slab, uh, warm, steps):
        """W eager warm-up passes, capture, `steps` replays between device events; returns (ms, launches, state)"""
        grid = slab.grid
        slab.begin(0.0, 0.0)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            grid.use_current_stream()
            for _ in range(warm):
                slab.step(uh)
        sync_all()
        per_graph = 0
        if use_graph:
            before = grid.launch_count
            slab.capture(uh)
            per_graph = grid.launch_count - before            # kernels in one captured pass
            one_step = slab.replay
        else:
            grid.use_current_stream()
            one_step = lambda: slab.step(uh)
        sync_all()
        before = grid.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(steps):
            one_step()
        ev1.record()
        sync_all()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = per_graph * steps if use_graph else grid.launch_count - before
        return ms, launches, slab.state()

    def drop(slab):
        torch.cuda.synchronize()
        for obj in (slab, slab.alg):
            if hasattr(obj, "_graph"):
                del obj._graph
        torch.cuda.synchronize()
        slab.close()

    # ---- the headline: weak scaling, every GPU owns a 512^3 copy of the unit-cube problem -----------------------------
    n_global = (N_SIDE, N_SIDE, N_SIDE * world)
    h = (1.0 / N_SIDE,) * 3
    slab = make_slab(n_global, h)
    grid = slab.grid
    uh = slab.color()
    unit = ops.CartesianGrid((N_SIDE,) * 3, h=h, device=local_rank)
    u_unit = unit.init(ops.PROB_LEVEQUE)
    uh.zero_()
    slab.owned(uh).copy_(u_unit.view(N_SIDE, -1))              # sphere at (0.35, 0.35, 0.35) + k e_z in every slab
    del u_unit
    unit.close()
    cells_per_gpu = N_SIDE ** 3
    total_cells = cells_per_gpu * world

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, state = timed_run(slab, uh, W, K)
    if world > 1:
        ln = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    value = total_cells * K / (ms * 1e-3)

    # per-kernel device times of the same workload (separate eager passes: the event pairs perturb the headline slightly)
    grid.use_current_stream()
    vlib.check(L.vofx_profile_enable(grid._ctx, 1))
    KP = min(K, 50)
    for _ in range(KP):
        slab.step(uh)
    names = C.create_string_buffer(48 * 32)
    tot = (C.c_double * 32)()
    cnt = (C.c_int64 * 32)()
    nent = C.c_int(0)
    vlib.check(L.vofx_profile_read(grid._ctx, 32, names, tot, cnt, C.byref(nent)))
    vlib.check(L.vofx_profile_enable(grid._ctx, 0))
    sampler.stop()
    sync_all()
    kernels = {}
    for i in range(nent.value):
        nm = names.raw[48 * i:48 * (i + 1)].split(b"\0", 1)[0].decode()
        kernels[nm] = {"avg_ms": tot[i] / max(cnt[i], 1), "launches_per_step": cnt[i] / KP}
    step_ms_profiled = sum(k["avg_ms"] * k["launches_per_step"] for k in kernels.values())
    for k in kernels.values():
        k["share"] = (k["avg_ms"] * k["launches_per_step"]) / step_ms_profiled if step_ms_profiled > 0 else None
    mixed_per_rank = gather_ints(state.mixed_cells)

    # ---- end to end through the host-buffer entry points: H2D colour, one loop pass, D2H, every step ------------------
    if world == 1:
        host_u = torch.empty(grid.cells, dtype=torch.float64).pin_memory()
        host_u.copy_(uh)
        torch.cuda.synchronize()
        p = slab.alg.params
        KE = max(1, args.e2e_steps)
        vlib.check(L.vofx_step_host(grid._ctx, C.byref(p), C.c_void_p(host_u.data_ptr()), None))   # warm-up (allocates mirrors)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(KE):
            vlib.check(L.vofx_step_host(grid._ctx, C.byref(p), C.c_void_p(host_u.data_ptr()), None))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        vlib.check(L.vofx_host_transfer_bytes(grid._ctx, C.byref(h2d), C.byref(d2h)))
        e2e = {"value": total_cells * KE / wall, "unit": "cell-steps/s", "h2d_bytes_per_step": int(h2d.value),
               "d2h_bytes_per_step": int(d2h.value), "steps": KE, "ms_per_step": 1e3 * wall / KE,
               "api": "vofx_step_host: pinned host colour function in (whole array H2D, dense flags), updated in place on the host from "
                      "the (cell, value) pairs of the cells the step changed (sparse D2H) + loop state"}
        # the call example/vof.cc makes: Algorithm::operator()( uh, start, end ) (algorithm.hh:54-103) = the host colour function
        # up once, the loop passes on the device-resident copy, the colour function back for the data writer.  Reported next to
        # the per-step figure above (which copies the whole array every step, what an operator-by-operator driver pays).
        try:
            KA = 100
            t0 = time.perf_counter()
            vlib.check(L.vofx_color_upload(grid._ctx, C.c_void_p(host_u.data_ptr())))
            vlib.check(L.vofx_step_begin(grid._ctx, C.c_double(state.time), C.c_double(state.dt)))
            for _ in range(KA):
                vlib.check(L.vofx_step_resident(grid._ctx, C.byref(p)))
            vlib.check(L.vofx_color_download(grid._ctx, C.c_void_p(host_u.data_ptr()), None))
            torch.cuda.synchronize()
            wall_a = time.perf_counter() - t0
            e2e["algorithm_call"] = {"value": total_cells * KA / wall_a, "unit": "cell-steps/s", "passes_per_call": KA,
                                     "ms_per_step": 1e3 * wall_a / KA, "h2d_bytes_per_call": 8 * grid.cells, "d2h_bytes_per_call": 8 * grid.cells,
                                     "api": "vofx_color_upload (pinned host colour function, whole array) + vofx_step_begin + 100 x vofx_step_resident "
                                            "(eager launches, incremental flagging) + vofx_color_download, all inside the timed region: what "
                                            "Dune::VoFx::Algorithm::operator() does between two writes of the data writer"}
        except Exception as exc:
            e2e["algorithm_call"] = {"error": repr(exc)}
        # a driver that keeps the colour function on the host and declares what it changed there between two passes
        # (nothing, in a plain time loop): vofx_step_host_sparse uploads only declared cells and writes the changed cells back
        try:
            vlib.check(L.vofx_step_host_sparse(grid._ctx, C.byref(p), C.c_void_p(host_u.data_ptr()), None, None, 0))   # first call: whole array up
            torch.cuda.synchronize()
            KS = 50
            t0 = time.perf_counter()
            for _ in range(KS):
                vlib.check(L.vofx_step_host_sparse(grid._ctx, C.byref(p), C.c_void_p(host_u.data_ptr()), None, None, 0))
            torch.cuda.synchronize()
            wall_s = time.perf_counter() - t0
            vlib.check(L.vofx_host_transfer_bytes(grid._ctx, C.byref(h2d), C.byref(d2h)))
            e2e["host_sparse"] = {"value": total_cells * KS / wall_s, "unit": "cell-steps/s", "steps": KS, "ms_per_step": 1e3 * wall_s / KS,
                                  "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value),
                                  "api": "vofx_step_host_sparse with an empty list of touched cells: the host colour function is current after "
                                         "every pass (changed (cell, value) pairs device -> host), nothing goes host -> device because the "
                                         "caller declares that it changed nothing"}
        except Exception as exc:
            e2e["host_sparse"] = {"error": repr(exc)}
        drop(slab)
    else:
        # every rank: H2D of its owned slab from pinned host memory, one loop pass on a caller-owned device tensor (NCCL halo
        # exchange + MIN all-reduce: the host owns the colour function, so nothing may be kept in peer windows), the changed
        # (cell, value) pairs back into the host slab
        owned_now = slab.owned(uh).clone()                      # the window is freed with the context
        drop(slab)
        slab_e = make_slab(n_global, h, tr="nccl")
        ge = slab_e.grid
        ue = ge.empty("color")
        H = slab_e.exchanger.halo
        owned = slab_e.owned(ue)
        host_u = torch.empty(owned.numel(), dtype=torch.float64).pin_memory()
        host_u.copy_(owned_now.reshape(-1))
        del owned_now
        ge.use_current_stream()
        vlib.check(L.vofx_track_changes(ge._ctx, 1))
        vlib.check(L.vofx_set_incremental_flagging(ge._ctx, 0))   # the host owns the colour function here: dense flags every pass
        slab_e.begin(state.time, state.dt)
        KE = max(1, args.e2e_steps)
        nchg = C.c_int64(0)

        def e2e_step():
            owned.copy_(host_u.view_as(owned), non_blocking=True)
            slab_e.step(ue)
            vlib.check(L.vofx_changed_fetch(ge._ctx, C.c_void_p(host_u.data_ptr()), C.c_int64(-H * slab_e.layer_cells), C.byref(nchg)))

        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(KE):
            e2e_step()
        sync_all()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_cells * KE / wall, "unit": "cell-steps/s", "h2d_bytes_per_step": 8 * cells_per_gpu * world,
               "d2h_bytes_per_step": (12 * int(nchg.value) + 64) * world, "steps": KE, "ms_per_step": 1e3 * wall / KE,
               "api": "per rank: pinned host slab -> device, SlabAlgorithm(transport='nccl').step (halo exchange + MIN reduction over NCCL, dense "
                      "flags), vofx_changed_fetch writes the changed (cell, value) pairs back into the host slab; d2h bytes estimated "
                      "from rank 0's count"}
        drop(slab_e)

    # ---- N > 1: the configs that are ONE problem cut into slabs, with the band balanced over the ranks ---------------------
    workloads = {}
    multirank_parity = None
    if world > 1 and not args.no_extra_workloads:
        def one_problem(tag, n_glob, hh, recon=ops.RECON_YOUNGS, steps=None):
            """one sphere in n_glob: even split -> per-layer cost from the initial flags -> cost-weighted split -> timed run"""
            steps = steps or args.extra_steps
            res = {"grid": list(n_glob)}
            out = {}
            for mode in ("even", "balanced"):
                layers = None
                if mode == "balanced":
                    layers = out["_counts"]
                    if layers == out["even"]["layers"]:
                        res["balanced"] = dict(res["even"], note="the cost-weighted split equals the even one")
                        break
                s2 = make_slab(n_glob, hh, layers=layers, recon=recon)
                u2 = s2.color()
                s2.grid.init(ops.PROB_LEVEQUE, out=u2)          # analytic initial data in global coordinates, halos included
                torch.cuda.synchronize()
                if mode == "even":
                    ops.FlagOperator(s2.grid, EPS)(u2, s2.alg.flags)
                    costs = s2.global_layer_costs()
                    out["_counts"] = split_by_cost(costs, world, min_layers=max(s2.exchanger.halo, 1))
                ms2, _, st2 = timed_run(s2, u2, W, steps)
                cells = 1
                for x in n_glob:
                    cells *= x
                entry = {"layers": list(s2.layer_counts), "ms_per_step": ms2 / steps, "steps": steps,
                         "value": cells * steps / (ms2 * 1e-3), "unit": "cell-steps/s",
                         "value_per_gpu": cells * steps / (ms2 * 1e-3) / world,
                         "mixed_cells_per_rank": gather_ints(st2.mixed_cells), "time": st2.time}
                out[mode] = entry
                res[mode] = entry
                drop(s2)
            return res

        if N_SIDE % world == 0 and N_SIDE // world >= 4:
            workloads["c3_strong"] = dict(one_problem("c3_strong", (N_SIDE,) * 3, (1.0 / N_SIDE,) * 3),
                                          workload="configs[2] strong scaling: ONE 512^3 LeVeque sphere cut into z-slabs", scaling="strong")
        if 256 % world == 0:
            # fp64-bound: the band decides everything, so the cost-weighted split is the one that matters
            workloads["c4"] = dict(one_problem("c4", (256,) * 3, (1.0 / 256,) * 3, recon=ops.RECON_SWARTZ, steps=20),
                                   workload="configs[3]: ONE LeVeque sphere on 256^3, iterative Swartz reconstruction (maxIterations 10), z-slabs",
                                   scaling="strong", one_gpu_reference="profiles/r2_configs.jsonl (tools/bench_configs.py, c4)")
        if world == 8:
            n5 = 2 * N_SIDE
            workloads["c5"] = dict(one_problem("c5", (n5, n5, n5), (1.0 / n5,) * 3),
                                   workload="configs[4]: ONE LeVeque sphere on 1024^3 (1024 x 1024 x 128 per GPU on the even split)", scaling="weak vs 1 GPU x 512^3")

        # bit parity of the slab run with the serial run on a case whose sphere crosses the slab boundaries (outside any timed region)
        n_small = (48, 40, 64)
        g1 = ops.CartesianGrid(n_small, device=local_rank)
        g1.set_velocity_builtin(ops.PROB_LEVEQUE)
        u1 = g1.init(ops.PROB_LEVEQUE)
        u0 = u1.clone()
        st1 = ops.Algorithm(g1, CFL, EPS, reconstruction_kind=ops.RECON_YOUNGS).run_passes(u1, 10)
        sp = SlabAlgorithm(n_small, rank, world, CFL, EPS, ops.RECON_YOUNGS, device=local_rank, transport=transport)
        sp.set_velocity_builtin(ops.PROB_LEVEQUE)
        up = sp.color()
        up.zero_()
        lo = sum(sp.layer_counts[:rank])
        layer = u0.numel() // n_small[-1]
        sp.owned(up).copy_(u0.view(n_small[-1], layer)[lo:lo + sp.layer_counts[rank]])
        sp.begin(0.0, 0.0)
        for _ in range(10):
            sp.step(up)
        stp = sp.state()
        torch.cuda.synchronize()
        same = torch.equal(sp.owned(up).contiguous().view(torch.int64), u1.view(n_small[-1], layer)[lo:lo + sp.layer_counts[rank]].contiguous().view(torch.int64))
        ok = torch.tensor([1 if (same and stp.time == st1.time and stp.dt == st1.dt) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        multirank_parity = {"case": "48 x 40 x 64 LeVeque sphere, Young's, 10 passes, z-slabs vs one context", "transport": sp.transport,
                            "layers": list(sp.layer_counts), "bit_identical": bool(int(ok.item()) == 1)}
        sp.close()
        g1.close()

    def teardown():
        # regular interpreter exit (no os._exit): the driver's exit hook records the libraries this process mapped
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return 0

    peak, peak_src = measured_peaks()
    cs_per_gpu = value / world
    achieved = cs_per_gpu * SURVEY_BYTES_PER_CELL_STEP / 1e9
    step_ms = ms / K
    sub = {}
    for nm, k in kernels.items():
        sub[nm] = {"ms_per_step": k["avg_ms"] * k["launches_per_step"], "launches_per_step": k["launches_per_step"], "share_of_profiled_step": k["share"]}
    dense = kernels.get("k_flag_compact")
    if dense:
        dense_ms = dense["avg_ms"] * dense["launches_per_step"]
        sub["k_flag_compact"].update({"bound": "hbm", "algorithmic_bytes_per_cell": FLAG_KERNEL_BYTES_PER_CELL,
                                      "note": "dense flag kernel (k_flag_rows); with incremental flagging it only runs on the halo layers"})
        if not incremental and world == 1:
            sub["k_flag_compact"]["achieved_gbs"] = FLAG_KERNEL_BYTES_PER_CELL * cells_per_gpu / (dense_ms * 1e-3) / 1e9
            sub["k_flag_compact"]["frac_of_peak"] = sub["k_flag_compact"]["achieved_gbs"] / peak
    # fp64-pipe utilisation of the band kernels: from the committed ncu captures of this workload (not measured in this run)
    try:
        summary = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")))
        for nm, key in (("k_youngs", "k_youngs_coop3"), ("k_flux", "k_flux_cell3")):
            if nm in sub and key in summary and summary[key].get("fp64_pipe_active_pct") is not None:
                sub[nm]["bound"] = "fp64 latency"
                sub[nm]["fp64_pipe_frac"] = summary[key]["fp64_pipe_active_pct"] / 100.0
                sub[nm]["fp64_pipe_source"] = "profiles/r2_ncu_summary.json (static: ncu --set full capture of this workload)"
    except Exception:
        pass
    # DRAM traffic of one step from the committed ncu captures of this workload (static, per launch; the band kernels run as
    # two chains, a capture is one chain): far below the algorithmic 17 B/cell because the dense passes are gone
    traffic, traffic_note = None, "not captured in this run; per-kernel dram bytes of the ncu captures are under profiles/"
    try:
        summary = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")))
        per_step = {"k_flag_segments": 1, "k_youngs_coop3": 2, "k_locate_root3": 2, "k_flux_cell3": 2, "k_update": 1}
        if world == 1 and incremental and all(k in summary for k in per_step):
            traffic = float(sum(summary[k]["dram_bytes_total"] * n for k, n in per_step.items()))
            traffic_note = ("static: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full captures of this workload "
                            "(profiles/r2_ncu_summary.json), summed over the launches of one step (two band chains); not measured in this run")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "whole step (SURVEY.md 8(d): cell-steps/s per GPU x 17 B against the measured HBM peak)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "algorithmic_bytes_per_cell_step": SURVEY_BYTES_PER_CELL_STEP,
                "algorithmic_bytes_per_step_per_gpu": SURVEY_BYTES_PER_CELL_STEP * cells_per_gpu,
                "peak_source": peak_src, "ms_per_step": step_ms,
                "note": "the step moves far fewer bytes than 17 B/cell: the u' write is an in-place band update and, with incremental "
                        "flagging, the dense u read / flag write shrinks to the band's neighbourhood; what bounds the step is the fp64 "
                        "latency of the band kernels (sub-entries)",
                "kernels": sub}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            _, cpu_baseline = cpu_reference_run(passes=4, warm=1)
        except Exception as exc:   # the checker must never take the bench down
            cpu_baseline = {"value": None, "error": repr(exc)}

    line = {
        "metric": "VoF cell-steps/sec (reconstruct+advect)", "value": value, "unit": "cell-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME,
                   "grid": list(n_global), "cells_per_gpu": cells_per_gpu, "decomposition": f"z-slabs x{world}",
                   "transport": (slab.transport if world > 1 else None),
                   "flagging": "incremental (dense first pass, then re-flag of the band's two-step neighbourhood)" if incremental else "dense every step",
                   "l2": "inputs larger than L2 (colour function 1.07 GB per GPU vs 126 MB L2)",
                   "mixed_cells_last_step_per_rank": mixed_per_rank, "time": state.time, "dt": state.dt},
        "clocks": sampler.report(),
        "e2e": e2e,
        "gpu_launches": launches,
        "kernels": kernels,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        line["workloads"] = workloads
        line["multirank_parity"] = multirank_parity
    emit(json.dumps(line))
    teardown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
