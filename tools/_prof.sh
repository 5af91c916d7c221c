set -e
tools/erk_bench_cur 1048576 3
ncu --set full --clock-control none --import-source on -k regex:erk_fused -s 1 -c 1 -o gpurun_out/erk_cur tools/erk_bench_cur 1048576 1 > gpurun_out/ncu_cur.log 2>&1
tail -2 gpurun_out/ncu_cur.log
