set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1_j.json 2> gpurun_out/bench_r1_j.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1_j.json 2> gpurun_out/bench_ref_r1_j.err
timeout 900 python tools/bench_configs.py --configs c1,c2,c4 > gpurun_out/configs_r1_j.jsonl 2> gpurun_out/configs_r1_j.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g.log 2>&1
for k in k_flag_rows k_youngs_coop3 k_flux_cell3 k_update k_locate_root3; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/r1j_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
