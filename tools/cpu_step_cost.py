"""CPU cost of enqueuing one multi-GPU step (run under torchrun with 2 ranks): is the loop launch-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import dune_vof_b200.operators as ops
from dune_vof_b200.decomposition import SlabAlgorithm
N = 512
for overlap in (False, True):
    slab = SlabAlgorithm((N, N, N * world), rank, world, 0.5, 1e-6, ops.RECON_YOUNGS, h=(1.0 / N,) * 3, device=rank, overlap=overlap)
    g = slab.grid
    g.set_velocity_builtin(ops.PROB_LEVEQUE); g.use_current_stream()
    unit = ops.CartesianGrid((N,) * 3, h=(1.0 / N,) * 3, device=rank)
    u1 = unit.init(ops.PROB_LEVEQUE)
    uh = g.empty("color"); slab.owned(uh).copy_(u1.view(N, -1)); del u1; unit.close()
    slab.begin(0.0, 0.0)
    for _ in range(5):
        slab.step(uh)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        slab.step(uh)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        print(f"overlap={overlap}: CPU enqueue {1e3*(t1-t0)/100:.3f} ms/step, total {1e3*(t2-t0)/100:.3f} ms/step", flush=True)
    g.close()
dist.barrier(); dist.destroy_process_group()
