set -x
O=gpurun_out/r2l
mkdir -p $O
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do timeout 200 tests/adaptor/test_adaptor_parallel > $O/par_$i.log 2>&1; echo "run $i rc=$?"; grep -c "^ok" $O/par_$i.log; grep "rank\|FAIL" $O/par_$i.log | head -5; done
