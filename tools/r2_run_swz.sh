# batched Swartz with the per-iteration task sort: parity tests, then C4 with and without the sort on the same box
O=gpurun_out/r2swz
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu -k "swartz or Swartz or c4" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in sort nosort; do
  if [ $v = nosort ]; then export VOFX_SWZ_NO_SORT=1; else unset VOFX_SWZ_NO_SORT; fi
  timeout 200 python tools/bench_configs.py --configs c4 --no-cpu > $O/c4_$v.jsonl 2> $O/c4_$v.err
  python -c "
import json
d=json.loads(open('$O/c4_$v.jsonl').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],4), d['kernel_avg_ms'].get('k_swartz'))"
done
