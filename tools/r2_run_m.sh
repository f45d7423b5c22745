set -x
O=gpurun_out/r2m
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_distributed_local.py tests/test_gpu_baseline_sizes.py tests/test_gpu_adaptor.py -x -q -m gpu -k "incremental or local or c3 or c2 or adaptor" > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200.json 2> $O/bench200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20.json 2> $O/bench20.err
tail -3 $O/bench20.err
python - <<'PY'
import json
for f in ("bench200", "bench20"):
    try:
        d = json.load(open(f"gpurun_out/r2m/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
