# round 2, GPU call 24 (2 GPUs): rank-level table (GJB_STEP_LIGHT) -- R-rank bit-exactness at 2 ranks, bench with the rank-level and with the per-tile table
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29533 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -12 | tee gpurun_out/r2c24_dist_worker_2.log
GJB_STEP_LIGHT=0 GJB_TEST_STEP_ONLY=1 timeout 600 $TR --nproc-per-node 2 --master-port 29534 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -4 | tee gpurun_out/r2c24_dist_worker_2_fulltable.log
for L in 1 0; do
  GJB_STEP_LIGHT=$L timeout 400 $TR --nproc-per-node 2 --master-port 2954$L bench.py --gpus 2 --steps 20 --no-cpu-baseline > gpurun_out/r2c24_bench_g2_light$L.json 2> gpurun_out/r2c24_bench_g2_light$L.err; tail -2 gpurun_out/r2c24_bench_g2_light$L.err | cut -c1-300
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c24_bench_g2_light$L.json') if l.startswith('{')][-1]); print('LIGHT=$L N=2 global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['config']['logZ_check'][:30]))"
done
