"""Device-side timeline of one overlapped multi-GPU step (run under torchrun, 2 ranks)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
hp = os.environ.get("HP", "0") == "1"
if hp:
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank), pg_options=opts)
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import dune_vof_b200.operators as ops
from dune_vof_b200.decomposition import SlabAlgorithm
from dune_vof_b200.lib import check
N = 512
slab = SlabAlgorithm((N, N, N * world), rank, world, 0.5, 1e-6, ops.RECON_YOUNGS, h=(1.0 / N,) * 3, device=rank, overlap=True)
g, a = slab.grid, slab.alg
g.set_velocity_builtin(ops.PROB_LEVEQUE); g.use_current_stream()
unit = ops.CartesianGrid((N,) * 3, h=(1.0 / N,) * 3, device=rank)
u1 = unit.init(ops.PROB_LEVEQUE)
uh = g.empty("color"); slab.owned(uh).copy_(u1.view(N, -1)); del u1; unit.close()
slab.begin(0.0, 0.0)
for _ in range(10):
    slab.step(uh)
torch.cuda.synchronize(); dist.barrier()
main = torch.cuda.current_stream()
comm = torch.cuda.Stream(priority=-1) if hp else torch.cuda.Stream()
ev = {k: torch.cuda.Event(enable_timing=True) for k in "s x0 x1 p0 p1 r0 r1 p2 e".split()}
kap = C.c_void_p(0)
args = (g._ctx, C.byref(a.params), g.pc(uh), g.pf(a.flags), g.ph(a.reconstructions), kap)
acc = {}
for it in range(30):
    ev["s"].record(main)
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        ev["x0"].record(comm); slab.exchanger.exchange(slab.layers(uh)); ev["x1"].record(comm)
    check(g._L.vofx_step_phase(*args, 0)); ev["p0"].record(main)
    main.wait_stream(comm)
    check(g._L.vofx_step_phase(*args, 1)); ev["p1"].record(main)
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        ev["r0"].record(comm); dist.all_reduce(slab._dt_est, op=dist.ReduceOp.MIN); ev["r1"].record(comm)
    check(g._L.vofx_step_phase(*args, 2)); ev["p2"].record(main)
    main.wait_stream(comm)
    check(g._L.vofx_step_finish(g._ctx, C.byref(a.params))); ev["e"].record(main)
    torch.cuda.synchronize()
    for k in "x0 x1 p0 p1 r0 r1 p2 e".split():
        acc[k] = acc.get(k, 0.0) + ev["s"].elapsed_time(ev[k])
if rank == 0:
    print("hp=%d  ms since step start: " % hp + "  ".join(f"{k}={acc[k]/30:.3f}" for k in "x0 x1 p0 p1 r0 r1 p2 e".split()), flush=True)
dist.barrier(); dist.destroy_process_group()
