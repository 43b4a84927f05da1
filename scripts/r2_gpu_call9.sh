set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GJB_NVCC_EXTRA=-DGJB_TRACE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scratch/trace_step_dist.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -30 | tee gpurun_out/r2c9_trace_dist.txt
GJB_NVCC_EXTRA=-DGJB_TRACE GJB_STEP_TABLE=1 CUDA_VISIBLE_DEVICES=0 timeout 600 python scratch/trace_step.py --dim 1 2>&1 | tail -16 | tee gpurun_out/r2c9_trace_d1_table.txt
