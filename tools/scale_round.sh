# weak-scaling sweep on one box: bench.py at N = 1, 2, 4, 8 (run under `gpurun --gpus 8`)
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/w$N.err > gpurun_out/w$N.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/w$N.err > gpurun_out/w$N.json
  fi
  python -c "
import json; d=json.load(open('gpurun_out/w$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
