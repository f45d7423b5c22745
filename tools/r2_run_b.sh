# round 2: incremental flagging (segment marks): tests + bench + ncu of the flag kernels
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "incremental or c3 or c2" > $O/pytest_new.log 2>&1
tail -5 $O/pytest_new.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_incr.json 2> $O/bench_incr.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_incr20.json 2> $O/bench_incr20.err
python - <<'PY'
import json
for f in ("bench_incr", "bench_incr20"):
    try:
        d = json.load(open(f"gpurun_out/r2b/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
grep -E "k_mark_band|k_scan_marks|k_flag_segments|k_youngs|k_flux_cell|k_update|k_locate_root" $O/launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200- | tail -24
