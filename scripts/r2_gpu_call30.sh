# round 2, GPU call 30 (1 GPU): vector model kernels compiled for 4 CTAs per SM -- full suite again, the scoring-only shape, d = 32 line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2c30_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/bench_assess.py --dim 32 2>&1 | grep "^{" | tee gpurun_out/r2c30_assess_d32.json
timeout 300 python scripts/bench_assess.py --dim 8 --particles 16777216 2>&1 | grep "^{" | tee gpurun_out/r2c30_assess_d8.json
timeout 300 python bench.py --dim 32 --no-cpu-baseline --steps 10 2>/dev/null | grep "^{" | tee gpurun_out/r2c30_bench_d32.json | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | grep "^{" | tee gpurun_out/r2c30_bench_d1.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:model_kernel -c 2 --csv --log-file gpurun_out/r2c30_ncu_assess_d32.csv python scripts/bench_assess.py --dim 32 --iters 1 --warmup 0 > gpurun_out/ncu_assess.log 2>&1; tail -1 gpurun_out/ncu_assess.log | cut -c1-200
