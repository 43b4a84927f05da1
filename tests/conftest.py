import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # `unverified` used to auto-skip tests that had never run on a device; every one of them ran green on a B200 in
    # round 2 (profiles/r2_call1_gpu_tests.log), so the marker is now informational only
    config.addinivalue_line("markers", "unverified: written in round 1 after the GPU budget was spent; first device run in round 2 (green)")


@pytest.fixture(scope="session")
def device():
    import torch

    if os.environ.get("GJB_EMULATE") == "1" and not torch.cuda.is_available():
        # HOST-LOGIC dry run of the GPU tests on a CPU box: gjb_model_launch is emulated with the oracle
        # (tests/abi_emulator.py); tests that reach any other entry point fail.  Proves nothing about the kernels.
        import abi_emulator

        mp = pytest.MonkeyPatch()
        dev = abi_emulator.install(mp)
        yield dev
        mp.undo()
        return
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    yield torch.device("cuda", 0)
