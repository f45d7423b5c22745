# round-2 evidence on one B200: GPU suite, bench (200 / 20 steps, reference arm), BASELINE configs, launch list, ncu captures of
# the hot kernels of C3 and C4, fp64 op counters of one C4 pass.  TAG names the output directory under gpurun_out/.
set -x
TAG=${1:-r2final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1
tail -3 $O/pytest.log
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench200.json 2> $O/bench200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench20.json 2> $O/bench20.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python tools/bench_configs.py --configs c1,c2,c4 > $O/configs.jsonl 2> $O/configs.err
tail -2 $O/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
# the reports are tens of MB each and gpurun_out/ travels back only below 64 MiB: summarise on the box, keep three reports
for k in k_youngs_coop3 k_flux_cell3 k_flag_segments k_update k_locate_root3; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o $O/ncu_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_$k.log 2>&1
done
python tools/ncu_summary.py $O/ncu_ r2 $O
for k in k_youngs_coop3 k_flux_cell3; do
  python tools/ncu_opcodes.py $O/ncu_$k.ncu-rep > $O/opcodes_$k.txt 2>&1
done
rm -f $O/ncu_k_flag_segments.ncu-rep $O/ncu_k_update.ncu-rep $O/ncu_k_locate_root3.ncu-rep
cat > /tmp/c4pass.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
import dune_vof_b200.operators as ops
n = (256, 256, 256)
g = ops.CartesianGrid(n); g.set_velocity_builtin(ops.PROB_LEVEQUE); u = g.init(ops.PROB_LEVEQUE)
a = ops.Algorithm(g, 0.5, 1e-6, ops.RECON_SWARTZ, incremental=True)
a.begin(0.0, 0.0)
for _ in range(3):
    a.step(u)
torch.cuda.synchronize()
torch.cuda.profiler.start()
a.step(u)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("mixed", a.state().mixed_cells)
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $O/c4_pass_counters.csv python /tmp/c4pass.py > $O/c4_pass.log 2>&1
tail -2 $O/c4_pass.log
mkdir -p $O/c4
for k in k_swz_locate k_swz_centroid k_swz_normal; do
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o $O/c4/ncu_$k -f python /tmp/c4pass.py > $O/ncu_$k.log 2>&1
done
python tools/ncu_summary.py $O/c4/ncu_ r2 $O
python tools/ncu_opcodes.py $O/c4/ncu_k_swz_locate.ncu-rep > $O/opcodes_k_swz_locate.txt 2>&1
rm -rf $O/c4
du -sh $O
python - <<PY
import json
for f in ("bench200", "bench20"):
    try:
        d = json.load(open("$O/" + f + ".json"))
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k: round(v["avg_ms"] * v["launches_per_step"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
for line in open("$O/configs.jsonl"):
    try:
        d = json.loads(line)
        print(d["config"], round(d["ms_per_step"], 4), d["value"], d["roofline"].get("frac"), d["clocks"])
    except Exception as e:
        print("configs line failed", e)
PY
ls -la $O
