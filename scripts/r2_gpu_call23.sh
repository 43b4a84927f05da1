# round 2, GPU call 23 (8 GPUs): final code -- R-rank bit-exactness at 8 ranks, the bench at N = 8 / 4 / 2 (global default, islands inside),
# configs[3] on 8 and 1 GPU, configs[2] MH and configs[4] HMC sharded over 8 GPUs with the lane fingerprints of the 1-GPU run
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GJB_TEST_STEP_ONLY=1 GJB_TEST_N=20480 timeout 600 $TR --nproc-per-node 8 --master-port 29533 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -8 | tee gpurun_out/r2c23_dist_worker_8.log
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2954$N bench.py --gpus $N --steps 20 --no-cpu-baseline > gpurun_out/r2c23_bench_g$N.json 2> gpurun_out/r2c23_bench_g$N.err; tail -2 gpurun_out/r2c23_bench_g$N.err | cut -c1-300
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c23_bench_g$N.json') if l.startswith('{')][-1]); print('N=$N global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f value %.3e | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['islands']['value'], d['config']['logZ_check'][:30]))"
done
timeout 600 $TR --nproc-per-node 8 --master-port 29551 scripts/bench_configs.py hmm 2>&1 | grep "^{" | tee gpurun_out/r2c23_config3_hmm_8gpu.json | cut -c1-400
timeout 600 $TR --nproc-per-node 8 --master-port 29552 scripts/bench_configs.py mh 2>&1 | grep "^{" | tee gpurun_out/r2c23_config2_mh_8gpu.json | cut -c1-500
timeout 600 $TR --nproc-per-node 8 --master-port 29553 scripts/bench_configs.py hmc 2>&1 | grep "^{" | tee gpurun_out/r2c23_config4_hmc_8gpu.json | cut -c1-700
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/bench_configs.py hmm 2>&1 | grep "^{" | tee gpurun_out/r2c23_config3_hmm_1gpu.json | cut -c1-400
