# quick perf check: arithmetic self-test, 3 parity tests, bench at 20 and 200 steps
set -x
O=gpurun_out/r2e
mkdir -p $O
true
tail -3 $O/pytest_quick.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200.json 2> $O/bench200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20.json 2> $O/bench20.err
python - <<'PY'
import json
for f in ("bench200", "bench20"):
    try:
        d = json.load(open(f"gpurun_out/r2e/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
