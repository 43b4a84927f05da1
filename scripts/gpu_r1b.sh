set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d1_v1.json 2> gpurun_out/bench_d1_v1.err; tail -3 gpurun_out/bench_d1_v1.err; cat gpurun_out/bench_d1_v1.json
timeout 300 python bench.py --steps 5 --warmup 3 --dim 32 --no-cpu-baseline > gpurun_out/bench_d32_v1.json 2> gpurun_out/bench_d32_v1.err; tail -3 gpurun_out/bench_d32_v1.err; cat gpurun_out/bench_d32_v1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_kernel -c 1 -f -o gpurun_out/prof_pf_kernel_d1 python scripts/profile_pf.py --dim 1 --T 20 > gpurun_out/ncu_pf_d1.log 2>&1; tail -2 gpurun_out/ncu_pf_d1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 2 -c 1 -f -o gpurun_out/prof_model_kernel_d1_v1 python scripts/profile_pf.py --dim 1 --T 6 --mode graph > gpurun_out/ncu_model_d1_v1.log 2>&1; tail -2 gpurun_out/ncu_model_d1_v1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 2 -c 1 -f -o gpurun_out/prof_model_kernel_d32_v1 python scripts/profile_pf.py --dim 32 --T 6 --mode graph > gpurun_out/ncu_model_d32_v1.log 2>&1; tail -2 gpurun_out/ncu_model_d32_v1.log
