# round 2, GPU call 22: ABI 16 (32 sites / 16 return leaves) -- full GPU suite, default bench line, the scoring-only streaming shape
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/r2c22_gpu_tests.log
timeout 300 python bench.py 2>&1 | grep "^{" | tee gpurun_out/r2c22_bench_d1.json | cut -c1-600
timeout 300 python scripts/bench_assess.py --dim 32 2>&1 | grep "^{" | tee gpurun_out/r2c22_assess_d32.json
timeout 300 python scripts/bench_assess.py --dim 1 --particles 33554432 2>&1 | grep "^{" | tee gpurun_out/r2c22_assess_d1.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:model_kernel -c 3 --csv --log-file gpurun_out/r2c22_ncu_assess_d32.csv python scripts/bench_assess.py --dim 32 --iters 1 --warmup 0 > gpurun_out/ncu_assess.log 2>&1; tail -2 gpurun_out/ncu_assess.log | cut -c1-300
