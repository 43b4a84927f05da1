# round 2, GPU call 31 (1 GPU): nested / unrolled Vmap, mask().vmap(), Vmap under an outer particle batch, the filter over a switching model
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_nested_gpu.py tests/test_switch_gpu.py tests/test_dist_vmap_gpu.py tests/test_zzz_unverified_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2c31_gpu_tests_dynamic.log
