# round 2, GPU call 35 (2 GPUs): conflict-free table build -- R-rank bit-exactness with 2048 and 4096 global tiles (8 / 16 records per
# thread of the table build: the shapes of 4 and 8 ranks at 1 M particles), default worker, bench at N = 2
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GJB_TEST_STEP_ONLY=1 GJB_TEST_N=2097152 timeout 300 $TR --nproc-per-node 2 --master-port 29533 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -4 | tee gpurun_out/r2c35_dist_worker_2_tiles2048.log
GJB_TEST_STEP_ONLY=1 GJB_TEST_N=4194304 timeout 300 $TR --nproc-per-node 2 --master-port 29534 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -4 | tee gpurun_out/r2c35_dist_worker_2_tiles4096.log
GJB_TEST_STEP_ONLY=1 timeout 300 $TR --nproc-per-node 2 --master-port 29535 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3 | tee gpurun_out/r2c35_dist_worker_2.log
timeout 400 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 20 --no-cpu-baseline > gpurun_out/r2c35_bench_g2.json 2> gpurun_out/r2c35_bench_g2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c35_bench_g2.json') if l.startswith('{')][-1]); print('N=2 global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['config']['logZ_check'][:30]))" || tail -3 gpurun_out/r2c35_bench_g2.err
