"""Cross-compile (sm_100a) every model library the GPU tests will ask for, HERE, so the GPU box only loads them.

Runs the GPU test files on the CPU under the IR emulation of tests/abi_emulator.py with GJB_PREBUILD=1: every
compile_ir() the tests reach also queues an nvcc build into genjax_b200/_lib/ (git-ignored, travels with gpurun).
Test outcomes of this pass are irrelevant (tests that need entry points without emulation fail early, after their
models were requested).  Usage: python scripts/prebuild_test_models.py [test files...]"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(path):
    env = dict(os.environ, GJB_EMULATE="1", GJB_RUN_UNVERIFIED="1", GJB_PREBUILD="1", CUDA_VISIBLE_DEVICES="",
               GJB_PREBUILD_JOBS="3")
    r = subprocess.run([sys.executable, "-m", "pytest", path, "-q", "-m", "gpu", "-p", "no:cacheprovider", "-x" if False else "-q"],
                       cwd=ROOT, env=env, capture_output=True, text=True)
    return path, r.stdout.strip().splitlines()[-1:] + [l for l in r.stdout.splitlines() if "[prebuild]" in l]


if __name__ == "__main__":
    files = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "tests", "test_*gpu*.py")))
    before = len(glob.glob(os.path.join(ROOT, "genjax_b200", "_lib", "model_*.so")))
    with ThreadPoolExecutor(max_workers=3) as ex:
        for path, tail in ex.map(one, files):
            print(os.path.basename(path), *tail)
    after = len(glob.glob(os.path.join(ROOT, "genjax_b200", "_lib", "model_*.so")))
    print(f"model libraries: {before} -> {after}")
