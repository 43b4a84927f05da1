# round 2, GPU call 36 (8 GPUs): the bench at N = 8 with the conflict-free table build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --steps 12 --no-cpu-baseline > gpurun_out/r2c36_bench_g8.json 2> gpurun_out/r2c36_bench_g8.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c36_bench_g8.json') if l.startswith('{')][-1]); print('N=8 global us/step %.2f value %.3e e2e %.3e | islands us/step %.2f value %.3e | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], d['islands']['ms_per_step']*10, d['islands']['value'], d['config']['logZ_check'][:30]))" || tail -3 gpurun_out/r2c36_bench_g8.err
