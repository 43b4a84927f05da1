"""The C-ABI boundary: every symbol include/genjax_b200.h declares is exported
by the built libraries, the ctypes mirrors have the C layout, and the entry
points validate arguments without a GPU (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "genjax_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_0-9]+\s*\*?\s*(gjb_[a-z_0-9]+)\s*\(", src, flags=re.M)
    core = [n for n in names if not n.startswith("gjb_model_")]
    model = [n for n in names if n.startswith("gjb_model_")]
    return core, model


@pytest.fixture(scope="module")
def libs():
    sys.path.insert(0, ROOT)
    from genjax_b200.runtime import build, cabi
    from genjax_b200 import workloads

    core = cabi.core()
    models = workloads.prebuild_all()
    return core, models, build


def test_header_declares_expected_entry_points():
    core, model = _declared()
    assert "gjb_resample_systematic" in core and "gjb_weight_mass" in core and "gjb_gather_rows" in core
    assert set(model) >= {"gjb_model_info", "gjb_model_launch", "gjb_model_mh_chain", "gjb_model_hmc_chain"}


def test_core_library_exports_every_declared_symbol(libs):
    core_lib, _, build = libs
    core, _ = _declared()
    raw = C.CDLL(str(build.LIB / "libgjb_core.so"))
    for name in core:
        assert hasattr(raw, name), f"libgjb_core.so does not export {name}"
    from genjax_b200.runtime import cabi

    assert set(cabi.CORE_PROTOTYPES) == set(core), "ctypes prototypes and the header disagree"
    version = int(re.search(r"#define GJB_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert core_lib.gjb_abi_version() == version
    from genjax_b200.runtime import cabi

    assert cabi.ABI_VERSION == version  # the ctypes binding mirrors the same header revision


def test_model_libraries_export_every_declared_symbol(libs):
    _, models, _ = libs
    _, model = _declared()
    from genjax_b200.runtime import cabi

    assert set(cabi.MODEL_PROTOTYPES) == set(model)
    for name, cm in models.items():
        raw = C.CDLL(str(cm.path))
        for sym in model:
            assert hasattr(raw, sym), f"{name}: {cm.path.name} does not export {sym}"
        assert cm.info["name"] and len(cm.info["sites"]) >= 1


def test_ctypes_struct_layout_matches_c(tmp_path):
    """Compile a probe with gcc against the header and compare sizeof / offsetof."""
    from genjax_b200.runtime import cabi

    structs = {"gjb_model_args": cabi.ModelArgs, "gjb_resample_args": cabi.ResampleArgs, "gjb_chain_args": cabi.ChainArgs}
    if hasattr(cabi, "PfArgs"):
        structs["gjb_pf_args"] = cabi.PfArgs
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "genjax_b200.h"', "int main(void) {"]
    for cname, st in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, st in structs.items():
        assert int(out[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_argument_validation_without_gpu(libs):
    core, models, _ = libs
    assert core.gjb_resample_workspace_bytes(-1) == -1
    assert core.gjb_resample_workspace_bytes(0) == 8
    assert core.gjb_resample_workspace_bytes(2049) == 16
    assert core.gjb_wmax_reset(None, None) == -1  # GJB_E_ARG
    assert core.gjb_weight_max(None, 10, None, None) == -1
    assert core.gjb_weight_mass(None, 10, None, None, None, None) == -1
    assert core.gjb_gather_rows(None, None, None, 4, 4, None) == -1
    assert core.gjb_resample_systematic(None, None) == -1
    assert core.gjb_philox_fill(0, 0, 0, 0, 0, -5, None, None) == -1
    cm = next(iter(models.values()))
    assert cm.lib.gjb_model_launch(None, None) == -1


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import genjax_b200 as gj
    from genjax_b200.runtime.cabi import GjbError
    from genjax_b200.workloads import lgssm_step

    with pytest.raises(GjbError, match="no CPU fallback"):
        lgssm_step.simulate(gj.key(0), (0.0,))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "genjax_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
