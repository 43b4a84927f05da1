# round 2, GPU call 20: chain kernels (configs[2] MH, configs[4] HMC one GPU's share) -- timing after the block-size change, ncu pipe utilisation
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/bench_configs.py mh 2>&1 | grep "^{" | tee gpurun_out/r2c20_config2_mh_1gpu.json | cut -c1-400
timeout 300 python scripts/bench_configs.py hmc 2>&1 | grep "^{" | tee gpurun_out/r2c20_config4_hmc_1gpu.json | cut -c1-600
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:mh_chain_kernel -c 1 --csv --log-file gpurun_out/r2c20_ncu_mh_chain_kernel.csv python scripts/bench_configs.py mh > gpurun_out/ncu_mh.log 2>&1; tail -1 gpurun_out/ncu_mh.log | cut -c1-200
timeout 300 ncu --metrics $M --clock-control none -k regex:hmc_chain_kernel -c 1 --csv --log-file gpurun_out/r2c20_ncu_hmc_chain_kernel.csv python scripts/bench_configs.py hmc > gpurun_out/ncu_hmc.log 2>&1; tail -1 gpurun_out/ncu_hmc.log | cut -c1-200
GJB_HMC_CHAINS=65536 timeout 300 python scripts/bench_configs.py hmc 2>&1 | grep "^{" | tee gpurun_out/r2c20_config4_hmc_64k_1gpu.json | cut -c1-600
