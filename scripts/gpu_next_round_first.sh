# First GPU call of the next round: (1) the verified suite, (2) everything that was written after round 1's GPU
# budget was spent (tests marked `unverified`: get_subtrace, Scan, Vmap / repeat, the reference's static-language and
# request-composition tests), (3) smoke + the default bench line.  Un-mark what passes.
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
GJB_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests -m gpu -q -k "unverified or zzz" -rA 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed" | tee gpurun_out/unverified.log | tail -60
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/next_bench_d1.json 2> gpurun_out/next_bench_d1.err; tail -2 gpurun_out/next_bench_d1.err; cat gpurun_out/next_bench_d1.json
timeout 400 python bench.py --reference-max analytic --no-cpu-baseline > gpurun_out/next_bench_d1_analytic.json 2> gpurun_out/next_bench_d1_analytic.err; tail -2 gpurun_out/next_bench_d1_analytic.err; cat gpurun_out/next_bench_d1_analytic.json
timeout 400 python bench.py --reference-max analytic --single-pass --no-cpu-baseline > gpurun_out/next_bench_d1_single_pass.json 2> gpurun_out/next_bench_d1_single_pass.err; tail -2 gpurun_out/next_bench_d1_single_pass.err; cat gpurun_out/next_bench_d1_single_pass.json
# launch list + one full capture of the single-pass kernel (only meaningful if the unverified single-pass test passed above)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 120 --csv --log-file gpurun_out/next_launches_d1_single_pass.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --reference-max analytic --single-pass > gpurun_out/ncu_launches_sp.log 2>&1; tail -1 gpurun_out/ncu_launches_sp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel_static_pull -s 3 -c 1 -f -o gpurun_out/next_prof_model_kernel_static_pull_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph --reference-max analytic --single-pass > gpurun_out/ncu_sp.log 2>&1; tail -1 gpurun_out/ncu_sp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel_static_mass -s 3 -c 1 -f -o gpurun_out/next_prof_model_kernel_static_mass_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph --reference-max analytic > gpurun_out/ncu_sm.log 2>&1; tail -1 gpurun_out/ncu_sm.log
