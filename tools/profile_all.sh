# ncu captures of round 2 (run on the GPU box through gpurun); the big reports are summarised on the box and removed,
# because gpurun_out/ is limited to 64 MiB per call.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
summ() { python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1.ncu_summary.csv; python tools/ncu_blocks.py $O/$1.ncu-rep > $O/$1.blocks.txt 2>/dev/null; }
python tools/profile_step.py > $O/p_step.log 2>&1 && $NCU -k regex:erk_fused -s 1 -c 1 -o $O/r2_lorenz python tools/profile_step.py > $O/p_step_ncu.log 2>&1
summ r2_lorenz; rm -f $O/r2_lorenz.ncu-rep
python tools/profile_vdp.py > $O/p_vdp.log 2>&1 && $NCU -k regex:erk_fused -s 1 -c 1 -o $O/r2_vdp python tools/profile_vdp.py > $O/p_vdp_ncu.log 2>&1
summ r2_vdp; rm -f $O/r2_vdp.ncu-rep
python tools/profile_nbody.py 65536 > $O/p_nbody.log 2>&1 && $NCU -k regex:nbody_symplectic -s 1 -c 1 -o $O/r2_nbody python tools/profile_nbody.py 65536 > $O/p_nbody_ncu.log 2>&1
summ r2_nbody
python tools/profile_dense.py > $O/p_dense.log 2>&1 && $NCU -k regex:erk_fused -s 1 -c 1 -o $O/r2_dense python tools/profile_dense.py > $O/p_dense_ncu.log 2>&1
summ r2_dense; rm -f $O/r2_dense.ncu-rep
$NCU -k regex:"node_eval|node_locate" -c 2 -o $O/r2_dense_readers python tools/profile_dense.py > $O/p_dense2_ncu.log 2>&1
ncu -i $O/r2_dense_readers.ncu-rep --page raw --csv > $O/r2_dense_readers.raw.csv; rm -f $O/r2_dense_readers.ncu-rep
python tools/dgemm_timing.py 4096 8192 4096 > $O/p_dgemm.log 2>&1 && $NCU -k regex:"dgemm_dmma|d884gemm" -s 2 -c 2 -o $O/r2_dgemm python tools/dgemm_timing.py 4096 8192 4096 > $O/p_dgemm_ncu.log 2>&1
ncu -i $O/r2_dgemm.ncu-rep --page raw --csv > $O/r2_dgemm.raw.csv
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/p_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/p_bench_ncu.log 2>&1
du -sh $O; ls -la $O | tail -30
