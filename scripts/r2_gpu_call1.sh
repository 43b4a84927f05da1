# round 2, GPU call 1: the whole -m gpu suite INCLUDING the tests that never ran on a device, then the three step
# variants of the bench (default / analytic reference max / single pass) and ncu evidence for the two new kernels.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
GJB_RUN_UNVERIFIED=1 timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c1_tests.log; tail -40 gpurun_out/r2c1_tests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2c1_bench_d1.json 2> gpurun_out/r2c1_bench_d1.err; tail -2 gpurun_out/r2c1_bench_d1.err; cut -c1-400 gpurun_out/r2c1_bench_d1.json
timeout 400 python bench.py --reference-max analytic --no-cpu-baseline > gpurun_out/r2c1_bench_d1_analytic.json 2> gpurun_out/r2c1_bench_d1_analytic.err; tail -2 gpurun_out/r2c1_bench_d1_analytic.err; cut -c1-400 gpurun_out/r2c1_bench_d1_analytic.json
timeout 400 python bench.py --reference-max analytic --single-pass --no-cpu-baseline > gpurun_out/r2c1_bench_d1_single_pass.json 2> gpurun_out/r2c1_bench_d1_single_pass.err; tail -2 gpurun_out/r2c1_bench_d1_single_pass.err; cut -c1-400 gpurun_out/r2c1_bench_d1_single_pass.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 120 --csv --log-file gpurun_out/r2c1_launches_d1_single_pass.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --reference-max analytic --single-pass > gpurun_out/ncu_launches_sp.log 2>&1; tail -1 gpurun_out/ncu_launches_sp.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 120 --csv --log-file gpurun_out/r2c1_launches_d1_analytic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --reference-max analytic > gpurun_out/ncu_launches_an.log 2>&1; tail -1 gpurun_out/ncu_launches_an.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel_static_pull -s 3 -c 1 -f -o gpurun_out/r2c1_prof_model_kernel_static_pull_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph --reference-max analytic --single-pass > gpurun_out/ncu_sp.log 2>&1; tail -1 gpurun_out/ncu_sp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel_static_mass -s 3 -c 1 -f -o gpurun_out/r2c1_prof_model_kernel_static_mass_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph --reference-max analytic > gpurun_out/ncu_sm.log 2>&1; tail -1 gpurun_out/ncu_sm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resample_systematic_kernel -s 3 -c 1 -f -o gpurun_out/r2c1_prof_resample_systematic_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph --reference-max analytic > gpurun_out/ncu_rs.log 2>&1; tail -1 gpurun_out/ncu_rs.log
