set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d1_v4.json 2> gpurun_out/bench_d1_v4.err; tail -3 gpurun_out/bench_d1_v4.err; cat gpurun_out/bench_d1_v4.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode persistent > gpurun_out/bench_d1_v4p.json 2> gpurun_out/bench_d1_v4p.err; tail -3 gpurun_out/bench_d1_v4p.err; cat gpurun_out/bench_d1_v4p.json
timeout 300 python bench.py --steps 5 --warmup 3 --dim 32 --no-cpu-baseline > gpurun_out/bench_d32_v4.json 2> gpurun_out/bench_d32_v4.err; tail -3 gpurun_out/bench_d32_v4.err; cat gpurun_out/bench_d32_v4.json
