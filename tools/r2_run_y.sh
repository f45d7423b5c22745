# 2D clip / locate with cached level sets, cell volume and shared reciprocals: the whole GPU suite, then C1 / C2
O=gpurun_out/r2y
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python tools/bench_configs.py --configs c1,c2 --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<'PY'
import json
for l in open("gpurun_out/r2y/configs.jsonl"):
    d = json.loads(l); print(d["config"], round(d["ms_per_step"], 4), d["value"], d["roofline"].get("frac"), d["kernel_avg_ms"])
PY
