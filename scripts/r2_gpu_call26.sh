# round 2, GPU call 26 (8 GPUs): the three hand-off forms of the multi-GPU step at N = 8 / 4 / 2 -- rank-level table (GJB_STEP_LIGHT=1),
# per-tile table built by the last CTA (GJB_STEP_LIGHT=0), per-tile table built by the resident table kernel (GJB_STEP_TABLE_KERNEL=1);
# R-rank bit-exactness at 8 ranks with the rank-level table
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GJB_TEST_STEP_ONLY=1 GJB_TEST_N=20480 timeout 600 $TR --nproc-per-node 8 --master-port 29533 tests/dist_pf_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -6 | tee gpurun_out/r2c26_dist_worker_8_light.log
run() {  # name N env...
  name=$1; N=$2; shift 2
  env "$@" timeout 400 $TR --nproc-per-node $N --master-port 295$((40 + RANDOM % 50)) bench.py --gpus $N --steps 20 --no-cpu-baseline > gpurun_out/r2c26_bench_g${N}_$name.json 2> gpurun_out/r2c26_bench_g${N}_$name.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c26_bench_g${N}_$name.json') if l.startswith('{')][-1]); print('$name N=$N global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f value %.3e | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['islands']['value'], d['config']['logZ_check'][:30]))" || tail -3 gpurun_out/r2c26_bench_g${N}_$name.err
}
run light 8 GJB_STEP_LIGHT=1
run full 8 GJB_STEP_LIGHT=0
run tablekernel 8 GJB_STEP_TABLE_KERNEL=1
run light 4 GJB_STEP_LIGHT=1
run full 4 GJB_STEP_LIGHT=0
run light 2 GJB_STEP_LIGHT=1
run full 2 GJB_STEP_LIGHT=0
