# final state: smoke(), the whole GPU suite, the default bench line and the reference arm exactly as the driver runs them
O=gpurun_out/r2x
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference_default.json 2> $O/bench_reference_default.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_20.json 2> $O/bench_20.err
python - <<'PY'
import json
for f in ("bench_default", "bench_20", "bench_reference_default"):
    try:
        d = json.load(open(f"gpurun_out/r2x/{f}.json"))
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
