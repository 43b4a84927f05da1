# round 2, GPU call 32 (1 GPU): final build -- full -m gpu suite, smoke, default bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c32_gpu_tests.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2c32_gpu_tests.log | tail -12 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2c32_bench_d1.json 2> gpurun_out/r2c32_bench_d1.err; tail -2 gpurun_out/r2c32_bench_d1.err | cut -c1-200; cut -c1-330 gpurun_out/r2c32_bench_d1.json
