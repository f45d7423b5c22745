# round 2, state check after re-entry: GPU test-suite, bench (200 / 20 steps), reference arm, launch list, ncu of the hot kernels
set -x
O=gpurun_out/r2n
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench200.json 2> $O/bench200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20.json 2> $O/bench20.err
tail -3 $O/bench20.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
for k in k_youngs_coop3 k_flux_cell3 k_flag_segments k_update k_locate_root3; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o $O/ncu_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_$k.log 2>&1
done
python - <<'PY'
import json
for f in ("bench200", "bench20"):
    try:
        d = json.load(open(f"gpurun_out/r2n/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
ls -la $O
