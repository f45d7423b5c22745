# A/B on one box: libvofx variants (v0 baseline, v1 flux prefetch, v2 preloaded neighbour colours in the least squares, v3 both)
set -x
O=gpurun_out/r2s
mkdir -p $O
cp dune-vof_b200/libvofx.so /tmp/libvofx_current.so
for round in 1 2; do
for v in v0 v1 v2 v3; do
  cp variants/libvofx_$v.so dune-vof_b200/libvofx.so
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20_${v}_$round.json 2> $O/bench20_${v}_$round.err
done
done
for v in v0 v3; do
  cp variants/libvofx_$v.so dune-vof_b200/libvofx.so
  timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/bench200_$v.json 2> $O/bench200_$v.err
done
cp /tmp/libvofx_current.so dune-vof_b200/libvofx.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s/*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 4), {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items() if k in ("k_youngs", "k_flux")})
    except Exception as e:
        print(f, "failed", e)
PY
