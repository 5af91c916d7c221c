set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/profile_step.py > $O/p_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:erk_fused -s 1 -c 1 -f -o $O/r2_lorenz python tools/profile_step.py > $O/p_step_ncu.log 2>&1
python tools/profile_vdp.py > $O/p_vdp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:erk_fused -s 1 -c 1 -f -o $O/r2_vdp python tools/profile_vdp.py > $O/p_vdp_ncu.log 2>&1
python tools/profile_nbody.py 65536 > $O/p_nbody.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbody_symplectic -s 1 -c 1 -f -o $O/r2_nbody python tools/profile_nbody.py 65536 > $O/p_nbody_ncu.log 2>&1
python tools/profile_dense.py > $O/p_dense.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"erk_fused|node_" -s 1 -c 6 -f -o $O/r2_dense python tools/profile_dense.py > $O/p_dense_ncu.log 2>&1
python tools/dgemm_timing.py 4096 8192 4096 > $O/p_dgemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"dgemm_dmma|d884gemm" -s 2 -c 2 -f -o $O/r2_dgemm python tools/dgemm_timing.py 4096 8192 4096 > $O/p_dgemm_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $O/p_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $O/p_bench_ncu.log 2>&1
ls -la $O/*.ncu-rep
