# round 2, GPU call 15: full -m gpu suite on the final build flags (-fmad=false), smoke, default bench, compute-sanitizer passes
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c15_tests.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2c15_tests.log | tail -12 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2c15_bench_d1.json 2> gpurun_out/r2c15_bench_d1.err; tail -3 gpurun_out/r2c15_bench_d1.err; cut -c1-250 gpurun_out/r2c15_bench_d1.json
timeout 300 python bench.py --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c15_bench_d32.json 2> gpurun_out/r2c15_bench_d32.err; tail -3 gpurun_out/r2c15_bench_d32.err; cut -c1-200 gpurun_out/r2c15_bench_d32.json
SEL="tests/test_pf_step_gpu.py::test_step_filter_teacher_forced_vs_oracle tests/test_pf_step_gpu.py::test_tile_exponent_kernels_bit_exact"
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -q -x -p no:cacheprovider -k "2048 or 7-5 or 4100 or 2049" > gpurun_out/r2c15_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r2c15_sanitizer_$tool.log | cut -c1-200
done
