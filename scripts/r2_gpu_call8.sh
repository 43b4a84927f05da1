# round 2, GPU call 8 (2 GPUs): table-form tail after optimisation, graph vs eager, the 2-GPU bit-exactness worker, global bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" > gpurun_out/r2c8_$tag.json 2> gpurun_out/r2c8_$tag.err; tail -2 gpurun_out/r2c8_$tag.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c8_$tag.json') if l.startswith('{')][-1]); print('$tag'.ljust(20), 'us/step %.2f' % (d['ms_per_step']*10), 'kernel_us %.2f' % d['roofline']['kernel_us'], 'e2e %.3e' % d['e2e']['value'], 'value %.3e' % d['value'], d['config'].get('logZ_check','')[:40])"; }
run d1_default CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --mode step --no-cpu-baseline --steps 20
run d1_table CUDA_VISIBLE_DEVICES=0 GJB_STEP_TABLE=1 timeout 300 python bench.py --mode step --no-cpu-baseline --steps 20
GJB_TEST_STEP_ONLY=1 GJB_TEST_N=50000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_pf_worker.py 2>&1 | tail -12 | tee gpurun_out/r2c8_dist_worker.log
run g2_global timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --mode step --multi-gpu global --no-cpu-baseline --steps 20
run g2_islands timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --mode step --multi-gpu islands --no-cpu-baseline --steps 20
