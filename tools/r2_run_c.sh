set -x
O=gpurun_out/r2c
mkdir -p $O
for k in k_flag_segments k_scan_marks k_update; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 6 -c 1 -o $O/$k -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/ncu_$k.log 2>&1
done
ls -la $O
