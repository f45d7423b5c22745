set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_adaptor.py tests/test_gpu_parity.py -x -q -m gpu -k "adaptor or circle or c1_run or division" > $O/pytest.log 2>&1
tail -40 $O/pytest.log
for t in test_adaptor test_adaptor_parallel; do timeout 300 tests/adaptor/$t > $O/$t.log 2>&1; tail -25 $O/$t.log; done
