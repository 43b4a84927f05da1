# round 2, GPU call 21: dynamic structure (Switch / Mask / mix / or_else / Mask-ed constraints) and dist.repeat / dist.vmap on the device
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_switch_gpu.py tests/test_dist_vmap_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r2c21_gpu_tests_dynamic.log
