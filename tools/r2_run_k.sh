# whole GPU suite + smoke
set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 3000 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1
tail -6 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -2 $O/smoke.log
