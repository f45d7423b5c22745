# C4 after the scratch overlay of k_swz_normal: Swartz parity tests, then the C4 line
O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -k "swartz or Swartz or c4" > $O/pytest.log 2>&1
tail -3 $O/pytest.log
timeout 300 python tools/bench_configs.py --configs c4 --no-cpu > $O/c4.jsonl 2> $O/c4.err
python -c "
import json
d=json.loads(open('$O/c4.jsonl').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['kernel_avg_ms'])"
