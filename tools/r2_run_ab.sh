# A/B on one box: previous commit (variants/libvofx_prev.so) against the working tree's libvofx.so
O=gpurun_out/r2ab
mkdir -p $O
cp dune-vof_b200/libvofx.so /tmp/libvofx_new.so
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for round in 1 2; do
for v in prev new; do
  if [ $v = prev ]; then cp variants/libvofx_prev.so dune-vof_b200/libvofx.so; else cp /tmp/libvofx_new.so dune-vof_b200/libvofx.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/b20_${v}_$round.json 2> $O/b20_${v}_$round.err
  timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/b200_${v}_$round.json 2> $O/b200_${v}_$round.err
done
done
cp /tmp/libvofx_new.so dune-vof_b200/libvofx.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ab/b*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 4), {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items() if k in ("k_youngs", "k_flux", "k_cost_order")})
    except Exception as e:
        print(f, "failed", e)
PY
