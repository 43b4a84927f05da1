set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/final_bench_d1.json 2> gpurun_out/final_bench_d1.err; tail -2 gpurun_out/final_bench_d1.err; cat gpurun_out/final_bench_d1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -2 gpurun_out/final_bench_ref.err; cat gpurun_out/final_bench_ref.json
