# round 2, GPU call 13: the full -m gpu suite on the new defaults, smoke, default bench, d=32 streaming bench, ncu evidence
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c13_tests.log; tail -25 gpurun_out/r2c13_tests.log | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2c13_bench_d1.json 2> gpurun_out/r2c13_bench_d1.err; tail -3 gpurun_out/r2c13_bench_d1.err; cut -c1-250 gpurun_out/r2c13_bench_d1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c13_bench_ref.json 2> gpurun_out/r2c13_bench_ref.err; tail -3 gpurun_out/r2c13_bench_ref.err; cut -c1-300 gpurun_out/r2c13_bench_ref.json
timeout 300 python bench.py --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c13_bench_d32.json 2> gpurun_out/r2c13_bench_d32.err; tail -3 gpurun_out/r2c13_bench_d32.err; cut -c1-250 gpurun_out/r2c13_bench_d32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2c13_launches_d1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 3 -c 1 -f -o gpurun_out/r2c13_prof_pf_step_kernel_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 3 -c 1 -f -o gpurun_out/r2c13_prof_pf_step_kernel_d32 python scripts/profile_pf.py --dim 32 --T 10 --mode step --obs-sd 2.83 > gpurun_out/ncu_step32.log 2>&1; tail -2 gpurun_out/ncu_step32.log
