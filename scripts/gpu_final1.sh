set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/final_bench_d1.json 2> gpurun_out/final_bench_d1.err; tail -2 gpurun_out/final_bench_d1.err; cat gpurun_out/final_bench_d1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cat gpurun_out/final_bench_ref.json
timeout 300 python bench.py --steps 5 --warmup 3 --dim 32 --no-cpu-baseline > gpurun_out/final_bench_d32.json 2> gpurun_out/final_bench_d32.err; cat gpurun_out/final_bench_d32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 230 --csv --log-file gpurun_out/final_launches_d1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mass_resample_kernel -s 6 -c 1 -f -o gpurun_out/final_prof_mass_resample_kernel_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph > gpurun_out/ncu_mr.log 2>&1; tail -1 gpurun_out/ncu_mr.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 6 -c 1 -f -o gpurun_out/final_prof_model_kernel_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode graph > gpurun_out/ncu_mk.log 2>&1; tail -1 gpurun_out/ncu_mk.log
