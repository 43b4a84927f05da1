# round 2, GPU call 27 (8 GPUs): the bench at N = 8 / 4 / 2 with every timed step started from a barrier (default: per-tile table
# built by the last CTA), and the rank-level table at N = 8 / 2 for comparison under the same protocol
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # name N env...
  name=$1; N=$2; shift 2
  env "$@" timeout 400 $TR --nproc-per-node $N --master-port 295$((40 + RANDOM % 50)) bench.py --gpus $N --steps 20 --no-cpu-baseline > gpurun_out/r2c27_bench_g${N}_$name.json 2> gpurun_out/r2c27_bench_g${N}_$name.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c27_bench_g${N}_$name.json') if l.startswith('{')][-1]); print('$name N=$N global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f value %.3e | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['islands']['value'], d['config']['logZ_check'][:30]))" || tail -3 gpurun_out/r2c27_bench_g${N}_$name.err
}
run default 8 GJB_STEP_LIGHT=0
run default 4 GJB_STEP_LIGHT=0
run default 2 GJB_STEP_LIGHT=0
run light 8 GJB_STEP_LIGHT=1
run light 2 GJB_STEP_LIGHT=1
