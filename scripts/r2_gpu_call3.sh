# round 2, GPU call 3: the single-launch filter step (pf_step_kernel) -- parity tests, bench in step mode at d=1 and
# d=32, launch list and one full ncu capture of the kernel.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pf_step_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
timeout 300 python bench.py --mode step > gpurun_out/r2c3_bench_d1_step.json 2> gpurun_out/r2c3_bench_d1_step.err; tail -3 gpurun_out/r2c3_bench_d1_step.err; cut -c1-300 gpurun_out/r2c3_bench_d1_step.json
timeout 300 python bench.py --mode step --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c3_bench_d32_step.json 2> gpurun_out/r2c3_bench_d32_step.err; tail -3 gpurun_out/r2c3_bench_d32_step.err; cut -c1-300 gpurun_out/r2c3_bench_d32_step.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2c3_launches_d1_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode step > gpurun_out/ncu_launches_step.log 2>&1; tail -1 gpurun_out/ncu_launches_step.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 3 -c 1 -f -o gpurun_out/r2c3_prof_pf_step_kernel_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 3 -c 1 -f -o gpurun_out/r2c3_prof_pf_step_kernel_d32 python scripts/profile_pf.py --dim 32 --T 10 --mode step > gpurun_out/ncu_step32.log 2>&1; tail -2 gpurun_out/ncu_step32.log
