# band parts (VOFX_BAND_PARTS): parity of the whole GPU suite with 2 and 3 chains, then the bench per setting
set -x
O=gpurun_out/r2o
mkdir -p $O
VOFX_BAND_PARTS=2 timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_parts2.log 2>&1
tail -3 $O/pytest_parts2.log
VOFX_BAND_PARTS=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_incremental.py -x -q -m gpu > $O/pytest_parts3.log 2>&1
tail -3 $O/pytest_parts3.log
for P in 1 2 3 4; do
  VOFX_BAND_PARTS=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20_p$P.json 2> $O/bench20_p$P.err
  VOFX_BAND_PARTS=$P timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200_p$P.json 2> $O/bench200_p$P.err
done
python - <<'PY'
import json
for P in (1, 2, 3, 4):
    for f in ("bench20", "bench200"):
        try:
            d = json.load(open(f"gpurun_out/r2o/{f}_p{P}.json"))
            print(P, f, round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 4))
        except Exception as e:
            print(P, f, "failed", e)
PY
