# round 2, GPU call 4: phase timeline of pf_step_kernel (trace build)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GJB_NVCC_EXTRA=-DGJB_TRACE timeout 600 python scratch/trace_step.py --dim 1 2>&1 | tail -20 | tee gpurun_out/r2c4_trace_d1.txt
GJB_NVCC_EXTRA=-DGJB_TRACE timeout 600 python scratch/trace_step.py --dim 32 --obs-sd 2.83 2>&1 | tail -20 | tee gpurun_out/r2c4_trace_d32.txt
