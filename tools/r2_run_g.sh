# 2 GPUs: multirank bit parity (both transports), bench at N=2, bench at N=1 for reference
set -x
O=gpurun_out/r2g
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
true
true
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench2.json 2> $O/bench2.err
tail -3 $O/bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --transport nccl --no-extra-workloads > $O/bench2_nccl.json 2> $O/bench2_nccl.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err
python - <<'PY'
import json
for f in ("bench1", "bench2", "bench2_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r2g/{f}.json"))
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), d["value"], "e2e ms", round(d["e2e"]["ms_per_step"], 2), d.get("multirank_parity"))
        for k, v in (d.get("workloads") or {}).items():
            for mode in ("even", "balanced"):
                if mode in v:
                    print("   ", k, mode, v[mode]["layers"], round(v[mode]["ms_per_step"], 4), v[mode]["value_per_gpu"], v[mode]["mixed_cells_per_rank"])
    except Exception as e:
        print(f, "failed", e)
PY
