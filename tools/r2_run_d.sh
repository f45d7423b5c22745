# round 2: arithmetic self-test, whole GPU suite, bench (200 and 20 steps), launch list
set -x
O=gpurun_out/r2d
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "division" > $O/pytest_arith.log 2>&1
tail -3 $O/pytest_arith.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200.json 2> $O/bench200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20.json 2> $O/bench20.err
python - <<'PY'
import json
for f in ("bench200", "bench20"):
    try:
        d = json.load(open(f"gpurun_out/r2d/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1
tail -5 $O/pytest_all.log
