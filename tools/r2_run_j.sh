# C1 / C2 / C4 lines with clocks; fp64 op counters of one C4 pass; new tests
set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distributed_local.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "checkpoint or circle or local_slabs_match or c4" > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 900 python tools/bench_configs.py --configs c4 --no-cpu > $O/c4_first.jsonl 2> $O/c4_first.err
tail -2 $O/c4_first.err
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
wc -l $O/c4_pass_counters.csv
for k in k_swz_locate k_swz_normal; do
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o $O/$k -f python /tmp/c4pass.py > $O/ncu_$k.log 2>&1
done
ls -la $O
