# one 8-GPU box: bench.py at N = 1, 2, 4, 8 (weak scaling over peer windows + the one-problem workloads c3-strong, c4, c5 with
# even and cost-weighted slabs), the NCCL transport at 8 for comparison
O=gpurun_out/r2scale
mkdir -p $O
nvidia-smi topo -m > $O/topology.txt 2>&1
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> $O/w$N.err > $O/w$N.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline 2> $O/w$N.err > $O/w$N.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("$O/w$N.json"))
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("multirank_parity"))
    for k, w in (d.get("workloads") or {}).items():
        print("  ", k, {m: (round(w[m]["ms_per_step"], 4), w[m]["layers"]) for m in ("even", "balanced") if m in w})
except Exception as e:
    print("N=$N failed", e)
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --transport nccl --no-extra-workloads 2> $O/w8_nccl.err > $O/w8_nccl.json
python -c "
import json; d=json.load(open('$O/w8_nccl.json')); print('nccl', d['n_gpus'], d['value'], d['ms_per_step'])"
tail -3 $O/w8.err
