set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_pf_gpu.py tests/test_core_gpu.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d1_v5.json 2> gpurun_out/bench_d1_v5.err; tail -3 gpurun_out/bench_d1_v5.err; cat gpurun_out/bench_d1_v5.json
timeout 300 python bench.py --steps 5 --warmup 3 --dim 32 --no-cpu-baseline > gpurun_out/bench_d32_v5.json 2> gpurun_out/bench_d32_v5.err; tail -3 gpurun_out/bench_d32_v5.err; cat gpurun_out/bench_d32_v5.json
