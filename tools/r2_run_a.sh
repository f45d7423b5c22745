# round 2, first GPU pass: the new parity tests, the whole GPU suite, bench with dense and incremental flags
set -x
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt
free -g > $O/host_mem.txt; nproc >> $O/host_mem.txt
timeout 1500 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_baseline_sizes.py -x -q -m gpu > $O/pytest_new.log 2>&1
tail -5 $O/pytest_new.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_incr.json 2> $O/bench_incr.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --dense-flags > $O/bench_dense.json 2> $O/bench_dense.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_incr20.json 2> $O/bench_incr20.err
python - <<'PY'
import json
for f in ("bench_incr", "bench_dense", "bench_incr20"):
    try:
        d = json.load(open(f"gpurun_out/r2a/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_incremental.py --deselect tests/test_gpu_baseline_sizes.py > $O/pytest_rest.log 2>&1
tail -5 $O/pytest_rest.log
