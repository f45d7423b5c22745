# tunable sweep of the band kernels (none of them changes a result)
O=gpurun_out/r2sweep
mkdir -p $O
run () {
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2sweep/{sys.argv[1]}.json"))
    k = d["kernels"]
    print(sys.argv[1], round(d["ms_per_step"], 4), "youngs", round(k["k_youngs"]["avg_ms"], 4), "flux", round(k["k_flux"]["avg_ms"], 4), "update", round(k["k_update"]["avg_ms"], 4), flush=True)
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run base A=1
run y_bps20 VOFX_YOUNGS_BLOCKS_PER_SM=20
run y_bps30 VOFX_YOUNGS_BLOCKS_PER_SM=30
run y_bps60 VOFX_YOUNGS_BLOCKS_PER_SM=60
run y_bps80 VOFX_YOUNGS_BLOCKS_PER_SM=80
run y_l8 VOFX_LANES_PER_CELL=8
run f_bps4 VOFX_FLUX_BLOCKS_PER_SM=4
run f_bps12 VOFX_FLUX_BLOCKS_PER_SM=12
run f_bps16 VOFX_FLUX_BLOCKS_PER_SM=16
run f_bps32 VOFX_FLUX_BLOCKS_PER_SM=32
run f_l4 VOFX_FLUX_LANES=4
run f_l4_bps16 VOFX_FLUX_LANES=4 VOFX_FLUX_BLOCKS_PER_SM=16
run u_bps4 VOFX_UPDATE_BLOCKS_PER_SM=4
run u_bps16 VOFX_UPDATE_BLOCKS_PER_SM=16
run nocost VOFX_NO_COST_ORDER=1
