"""CUDA-graph capture of the multi-GPU step (kernels + NCCL ops): does replay hide the communication? (torchrun, 2 ranks)"""
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
    g.set_velocity_builtin(ops.PROB_LEVEQUE)
    unit = ops.CartesianGrid((N,) * 3, h=(1.0 / N,) * 3, device=rank)
    u1 = unit.init(ops.PROB_LEVEQUE)
    uh = g.empty("color"); slab.owned(uh).copy_(u1.view(N, -1)); del u1; unit.close()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g.use_current_stream()
        slab.begin(0.0, 0.0)
        for _ in range(5):
            slab.step(uh)
    torch.cuda.synchronize(); dist.barrier()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph, stream=side):
            g.use_current_stream()
            slab.step(uh)
    except Exception as exc:
        if rank == 0:
            print("capture failed:", repr(exc)[:300], flush=True)
        break
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"overlap={overlap}: graph replay {e0.elapsed_time(e1)/100:.4f} ms/step  state={slab.state()}", flush=True)
    g.close()
dist.barrier(); dist.destroy_process_group()
