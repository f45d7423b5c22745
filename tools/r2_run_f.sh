set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_distributed_local.py -x -q -m gpu > $O/pytest_local.log 2>&1
tail -15 $O/pytest_local.log
