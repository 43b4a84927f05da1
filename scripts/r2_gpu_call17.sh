# round 2, GPU call 17 (2 GPUs): small-footprint table kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c17_$1.json') if l.startswith('{')][-1]); i=d.get('islands') or {}; print('$1 us/step %.2f value %.3e e2e %.3e | islands us/step %.2f | %s' % (d['ms_per_step']*10, d['value'], d['e2e']['value'], i.get('ms_per_step',0)*10, d['config']['logZ_check'][:30]))"; }
CUDA_VISIBLE_DEVICES=0 GJB_STEP_TABLE=1 timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2c17_d1_table.json 2> gpurun_out/r2c17_d1_table.err; tail -1 gpurun_out/r2c17_d1_table.err | cut -c1-200; show d1_table
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 20 --no-cpu-baseline > gpurun_out/r2c17_g2.json 2> gpurun_out/r2c17_g2.err; tail -1 gpurun_out/r2c17_g2.err | cut -c1-200; show g2
GJB_PDL=0 timeout 300 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 20 --no-cpu-baseline > gpurun_out/r2c17_g2_nopdl.json 2> gpurun_out/r2c17_g2_nopdl.err; tail -1 gpurun_out/r2c17_g2_nopdl.err | cut -c1-200; show g2_nopdl
