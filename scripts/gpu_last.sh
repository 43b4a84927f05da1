set -x
cd $GRAFT_REPO_ROOT
timeout 400 python bench.py > gpurun_out/final_bench_d1.json 2> gpurun_out/final_bench_d1.err; tail -2 gpurun_out/final_bench_d1.err; cat gpurun_out/final_bench_d1.json
