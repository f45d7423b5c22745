# closed-form cell volume + two-phase least squares in the plane-constant kernel: parity of the GPU suite, bench per band-part setting
set -x
O=gpurun_out/r2p
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1
tail -3 $O/pytest.log
for P in 1 2; do
  VOFX_BAND_PARTS=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20_p$P.json 2> $O/bench20_p$P.err
  VOFX_BAND_PARTS=$P timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200_p$P.json 2> $O/bench200_p$P.err
done
python - <<'PY'
import json
for P in (1, 2):
    for f in ("bench20", "bench200"):
        try:
            d = json.load(open(f"gpurun_out/r2p/{f}_p{P}.json"))
            print(P, f, round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 4), {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
        except Exception as e:
            print(P, f, "failed", e)
PY
