# 8 GPUs: weak scaling (peer transport) at N = 1, 2, 4, 8 with c3-strong / c5 next to it, multirank parity
set -x
O=gpurun_out/r2h
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 > $O/bench$N.json 2> $O/bench$N.err
  tail -2 $O/bench$N.err
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 100 --warmup 5 --transport nccl --no-extra-workloads > $O/bench8_nccl.json 2> $O/bench8_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 tests/multirank_worker.py > $O/multirank8.log 2>&1
grep -c "bit-identical" $O/multirank8.log; grep "MULTIRANK\|MISMATCH" $O/multirank8.log
python - <<'PY'
import json
for f in ("bench1", "bench2", "bench4", "bench8", "bench8_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r2h/{f}.json"))
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), d["value"], "e2e ms", round(d["e2e"]["ms_per_step"], 2), (d.get("multirank_parity") or {}).get("bit_identical"))
        for k, v in (d.get("workloads") or {}).items():
            for mode in ("even", "balanced"):
                if mode in v:
                    print("   ", k, mode, v[mode]["layers"], round(v[mode]["ms_per_step"], 4), "%.4g" % v[mode]["value_per_gpu"], v[mode]["mixed_cells_per_rank"])
    except Exception as e:
        print(f, "failed", e)
PY
