import sys, time, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dune_vof_b200.operators as ops
from dune_vof_b200 import lib as vlib
N=512
g=ops.CartesianGrid((N,N,N)); g.set_velocity_builtin(ops.PROB_LEVEQUE); g.use_current_stream()
u=g.init(ops.PROB_LEVEQUE)
h=torch.empty(g.cells,dtype=torch.float64).pin_memory(); h.copy_(u); torch.cuda.synchronize()
for i in range(3):
    t0=time.perf_counter(); u.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    h.copy_(u, non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()
    print('H2D ms', 1e3*(t1-t0), 'D2H ms', 1e3*(t2-t1))
alg=ops.Algorithm(g,cfl=0.5,eps=1e-6,reconstruction_kind=ops.RECON_YOUNGS)
alg.begin(0.0,0.0)
L=g._L; p=alg.params
for i in range(5):
    t0=time.perf_counter(); vlib.check(L.vofx_step_host(g._ctx, C.byref(p), C.c_void_p(h.data_ptr()), None)); torch.cuda.synchronize(); t1=time.perf_counter()
    a,b=C.c_int64(0),C.c_int64(0); L.vofx_host_transfer_bytes(g._ctx,C.byref(a),C.byref(b))
    print('step_host ms', 1e3*(t1-t0), a.value, b.value)
