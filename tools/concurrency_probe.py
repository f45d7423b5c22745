"""Feasibility probe: can the dense flag pass (HBM-bound) run concurrently with the band kernels (fp64-latency-bound)?

One GPU, 512^3 LeVeque sphere after 20 passes.  Times (CUDA events)
  serial      : phase 0 (flags) ; phase 1 (Young's + flux) ; phase 2 (update)          -- the step as it is
  band        : phase 1 alone
  flag        : the dense flag pass alone (second context, separate outputs)
  concurrent  : phase 1 on stream A while the dense flag pass runs on stream B
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dune_vof_b200.operators as ops
from dune_vof_b200.lib import check

N = int(os.environ.get("N", "512"))
PRIO = os.environ.get("PRIO", "0") == "1"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
gA = ops.CartesianGrid((N,) * 3, device=0)
gB = ops.CartesianGrid((N,) * 3, device=0)
gA.set_velocity_builtin(ops.PROB_LEVEQUE)
uh = gA.init(ops.PROB_LEVEQUE)
alg = ops.Algorithm(gA, 0.5, 1e-6, ops.RECON_YOUNGS)
sA = torch.cuda.Stream(priority=-1) if PRIO else torch.cuda.Stream()
sB = torch.cuda.Stream()
flagsB = gB.empty("flags")
kap = C.c_void_p(0)
with torch.cuda.stream(sA):
    alg.begin(0.0, 0.0)
    for _ in range(20):
        alg.step(uh)
torch.cuda.synchronize()
args = (gA._ctx, C.byref(alg.params), gA.pc(uh), gA.pf(alg.flags), gA.ph(alg.reconstructions), kap)
with torch.cuda.stream(sB):
    gB.use_current_stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


res = {"serial": [], "band": [], "flag": [], "concurrent": [], "conc_band_end": [], "conc_flag_end": []}
for it in range(12):
    # serial step
    with torch.cuda.stream(sA):
        s, a, b, e = ev(), ev(), ev(), ev()
        s.record(sA)
        check(gA._L.vofx_step_phase(*args, 0)); a.record(sA)
        check(gA._L.vofx_step_phase(*args, 1)); b.record(sA)
        check(gA._L.vofx_step_phase(*args, 2))
        check(gA._L.vofx_step_finish(gA._ctx, C.byref(alg.params))); e.record(sA)
    torch.cuda.synchronize()
    res["serial"].append(s.elapsed_time(e)); res["band"].append(a.elapsed_time(b))
    # flag alone on B
    with torch.cuda.stream(sB):
        s, e = ev(), ev()
        s.record(sB)
        check(gB._L.vofx_flag(gB._ctx, gB.pc(uh), 1e-6, gB.pf(flagsB))); e.record(sB)
    torch.cuda.synchronize()
    res["flag"].append(s.elapsed_time(e))
    # concurrent: phase 0 first (valid list), then phase 1 on A || flag on B, then phase 2 on A
    with torch.cuda.stream(sA):
        check(gA._L.vofx_step_phase(*args, 0))
        s = ev(); s.record(sA)
    sB.wait_stream(sA)
    with torch.cuda.stream(sA):
        check(gA._L.vofx_step_phase(*args, 1)); ea = ev(); ea.record(sA)
    with torch.cuda.stream(sB):
        check(gB._L.vofx_flag(gB._ctx, gB.pc(uh), 1e-6, gB.pf(flagsB))); eb = ev(); eb.record(sB)
    sA.wait_stream(sB)
    with torch.cuda.stream(sA):
        e = ev(); e.record(sA)
        check(gA._L.vofx_step_phase(*args, 2))
        check(gA._L.vofx_step_finish(gA._ctx, C.byref(alg.params)))
    torch.cuda.synchronize()
    res["concurrent"].append(s.elapsed_time(e)); res["conc_band_end"].append(s.elapsed_time(ea)); res["conc_flag_end"].append(s.elapsed_time(eb))
med = lambda x: sorted(x)[len(x) // 2]
print("prio=%d  " % PRIO + "  ".join(f"{k}={med(v):.3f}ms" for k, v in res.items()), flush=True)
os._exit(0)
