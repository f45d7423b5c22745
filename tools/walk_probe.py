"""Per-step census of the plane-constant solve at 512^3 (needs a library built with the debug counters:
    make -C dune-vof_b200/csrc EXTRA=-DVOFX_COUNT_WALKS ; python tools/walk_probe.py ; then rebuild without the flag).
Prints how many located cells walked the topology on lane 0 (tied node heights), how many left the bracket loop at a node
(|V_k - V*| < 1e-14) and how many of those at the top node (nearly full cells: all 7 brackets are evaluated)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dune_vof_b200.operators as ops
from dune_vof_b200.lib import load
N = 512
g = ops.CartesianGrid((N,)*3, device=0)
g.set_velocity_builtin(ops.PROB_LEVEQUE)
uh = g.init(ops.PROB_LEVEQUE)
alg = ops.Algorithm(g, 0.5, 1e-6, ops.RECON_YOUNGS)
alg.begin(0.0, 0.0)
L = load()
out = (C.c_ulonglong * 8)()
L.vofx_debug_counters(out, 1)
for s in range(200):
    alg.step(uh)
    if s in (1, 20, 60, 120, 199):
        L.vofx_debug_counters(out, 1)
        print(s, "cells", out[0], "walking", out[1], "avg nU %.2f" % (out[2] / max(out[0], 1)), "f>=1", out[3], "f<=0", out[4], "exit at node", out[5], "of which top node", out[6], "first node", out[7], flush=True)
    elif s + 1 in (1, 20, 60, 120, 199):
        L.vofx_debug_counters(out, 1)
os._exit(0)
