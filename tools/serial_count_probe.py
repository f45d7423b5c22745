"""How many cells / clips per step leave the cooperative kernels for the serial ones (512^3 LeVeque sphere)?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import dune_vof_b200.operators as ops
N = 512
g = ops.CartesianGrid((N,)*3, device=0)
g.set_velocity_builtin(ops.PROB_LEVEQUE)
uh = g.init(ops.PROB_LEVEQUE)
alg = ops.Algorithm(g, 0.5, 1e-6, ops.RECON_YOUNGS)
alg.begin(0.0, 0.0)
out = []
kap = C.c_void_p(0)
args = (g._ctx, C.byref(alg.params), g.pc(uh), g.pf(alg.flags), g.ph(alg.reconstructions), kap)
from dune_vof_b200.lib import check
for s in range(60):
    check(g._L.vofx_step_phase(*args, 0))
    check(g._L.vofx_step_phase(*args, 1))
    torch.cuda.synchronize()
    cnt = (C.c_int64 * 4)()
    check(g._L.vofx_step_counters(g._ctx, cnt))
    check(g._L.vofx_step_phase(*args, 2))
    check(g._L.vofx_step_finish(g._ctx, C.byref(alg.params)))
    out.append(tuple(int(x) for x in cnt))
print("step: (mixed, serialClips, serialLocate, roots)")
for s in (0, 1, 2, 5, 10, 20, 40, 59):
    print(s, out[s])
os._exit(0)
