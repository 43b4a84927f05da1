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
    config.addinivalue_line(
        "markers",
        "unverified: GPU test of host-side code written after the round's GPU budget was spent; it has never run "
        "on a device, so it is skipped unless GJB_RUN_UNVERIFIED=1 (first GPU call of the next round)",
    )


def pytest_collection_modifyitems(config, items):
    if os.environ.get("GJB_RUN_UNVERIFIED") == "1":
        return
    skip = pytest.mark.skip(reason="never run on a GPU yet; set GJB_RUN_UNVERIFIED=1 (see DESIGN.md section 9)")
    for item in items:
        if "unverified" in item.keywords:
            item.add_marker(skip)


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
