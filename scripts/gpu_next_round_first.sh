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
