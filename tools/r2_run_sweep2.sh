# grid sizes of the two band kernels under the two-chain pipeline (each chain gets 1 / parts of the blocks)
O=gpurun_out/r2sweep2
mkdir -p $O
for YB in 40 60 90 120; do
for FB in 6 8 12 16; do
  VOFX_YOUNGS_BLOCKS_PER_SM=$YB VOFX_FLUX_BLOCKS_PER_SM=$FB timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/b_${YB}_${FB}.json 2> $O/b_${YB}_${FB}.err
  python -c "
import json; d=json.load(open('$O/b_${YB}_${FB}.json')); print($YB, $FB, round(d['ms_per_step'],4))"
done
done
