"""HOST-LOGIC dry run of GPU test files on CPU: GJB_EMULATE=1 makes the `device` fixture install
tests/abi_emulator.py (the oracle behind the real gjb_model_args / gjb_resample_args structures), so every host path
those tests walk -- argument binding, flags, choice maps, traces, SMC drivers, Scan, get_subtrace -- is exercised
without a device.  Proves nothing about the CUDA kernels (the -m gpu run on a B200 does); catches host regressions
between GPU runs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# the fused cooperative resampler and the multi-GPU links have no emulation; the Kalman comparisons and the 3M-particle
# case pass too but take a minute in NumPy
_PF = "not fused_mass_resample and not multi_gpu and not matches_kalman and not 3000000"
_MCMC = "not gmm"  # the mixture has no oracle sampler; everything else runs its generated chain kernels on the host


@pytest.mark.parametrize("files", [
    ["tests/test_core_gpu.py"],
    ["tests/test_gfi_gpu.py"],
    ["tests/test_zzz_unverified_gpu.py"],
    ["tests/test_zzz_static_reference_gpu.py"],
    ["tests/test_switch_gpu.py"],
    ["tests/test_dist_vmap_gpu.py"],
    ["tests/test_scan_nested_gpu.py"],
    # the filter, chain and core-kernel tests reach entry points that exist on the GPU only; these do not
    ["tests/test_pf_gpu.py", "-k", _PF],
    ["tests/test_zz_mv_normal_gpu.py"],
    ["tests/test_mcmc_gpu.py", "-k", _MCMC],
])
def test_gpu_tests_host_paths_under_emulation(files):
    # GJB_EMULATE_KERNELS=host: every model launch, chain launch and small filter runs the GENERATED CUDA source on the
    # host -- thread by thread for scalar-site models (tests/host_kernels.py), with real block semantics for vector-site
    # models, the persistent filter kernel and the kernels of libgjb_core (tests/simt_kernels.py)
    env = dict(os.environ, GJB_EMULATE="1", GJB_EMULATE_KERNELS="host", GJB_RUN_UNVERIFIED="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "pytest", *files, "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
