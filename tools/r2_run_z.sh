# vofx_step_host_sparse: its test, the host-buffer tests, and the bench line with the three end-to-end figures
O=gpurun_out/r2z
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b.json 2> $O/b.err; tail -3 $O/b.err
python -c "
import json; d=json.load(open('$O/b.json')); print(d['ms_per_step']); print(json.dumps({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='api'}) for k,v in d['e2e'].items() if k!='api'}, indent=1))"
