set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d1.json 2> gpurun_out/bench_d1.err; tail -3 gpurun_out/bench_d1.err; cat gpurun_out/bench_d1.json
for k in model_kernel weight_mass_kernel resample_systematic_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -f -o gpurun_out/prof_${k}_d1 python scripts/profile_pf.py --dim 1 > gpurun_out/ncu_${k}_d1.log 2>&1
  tail -2 gpurun_out/ncu_${k}_d1.log
done
ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 3 -c 2 -f -o gpurun_out/prof_model_kernel_d32 python scripts/profile_pf.py --dim 32 > gpurun_out/ncu_model_d32.log 2>&1
tail -2 gpurun_out/ncu_model_d32.log
ls -la gpurun_out
