# round 2, GPU call 5: step kernel on the table design (last-CTA prefix table, PDL) -- parity, bench with / without PDL
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pf_step_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
timeout 300 python bench.py --mode step --no-cpu-baseline > gpurun_out/r2c5_bench_d1_step_pdl.json 2> gpurun_out/r2c5_bench_d1_step_pdl.err; tail -3 gpurun_out/r2c5_bench_d1_step_pdl.err; cut -c1-300 gpurun_out/r2c5_bench_d1_step_pdl.json
GJB_PDL=0 timeout 300 python bench.py --mode step --no-cpu-baseline > gpurun_out/r2c5_bench_d1_step_nopdl.json 2> gpurun_out/r2c5_bench_d1_step_nopdl.err; tail -3 gpurun_out/r2c5_bench_d1_step_nopdl.err; cut -c1-300 gpurun_out/r2c5_bench_d1_step_nopdl.json
timeout 300 python bench.py --mode step --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c5_bench_d32_step.json 2> gpurun_out/r2c5_bench_d32_step.err; tail -3 gpurun_out/r2c5_bench_d32_step.err; cut -c1-300 gpurun_out/r2c5_bench_d32_step.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2c5_launches_d1_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode step > gpurun_out/ncu_launches_step.log 2>&1; tail -1 gpurun_out/ncu_launches_step.log
