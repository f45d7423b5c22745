"""Step time of PipelinedAlgorithm on one GPU at 512^3 for several chunk counts (graph replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dune_vof_b200.operators as ops
from dune_vof_b200.decomposition import PipelinedAlgorithm
N = int(os.environ.get("VOFX_BENCH_N", "512"))
n = (N, N, N)
base = ops.CartesianGrid(n)
u0 = base.init(ops.PROB_LEVEQUE)
base.close()
for chunks in [int(x) for x in os.environ.get("CHUNKS", "1,2,4,8").split(",")]:
    pipe = PipelinedAlgorithm(n, 0.5, 1e-6, ops.RECON_YOUNGS, chunks=chunks)
    pipe.set_velocity_builtin(ops.PROB_LEVEQUE)
    u = u0.clone()
    pipe.begin(0.0, 0.0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(5):
            pipe.step(u)
    torch.cuda.synchronize()
    pipe.capture(u)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        pipe.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"chunks={chunks}: {e0.elapsed_time(e1)/100:.4f} ms/step  {pipe.state()}", flush=True)
    pipe.close()
os._exit(0)
