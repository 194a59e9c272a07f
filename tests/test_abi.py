"""The C-ABI library: it builds for sm_100a, loads, and exports every symbol
include/sgk.h declares.  No compute calls here (no GPU needed)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sgk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgk_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from gridfast import _lib, build

    build.build()
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 30
    bound = {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), "libsgk.so does not export %s" % name
        assert name in bound, "gridfast._lib does not bind %s" % name
    assert lib.sgk_version() == 100


def test_header_cites_the_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "sgk.h")).read()
    for cite in ("train.py:51", "common/learn.py:61-85", "common/agents/value.py", "meters.py:66-108",
                 "ssrl/agents.py"):
        assert cite in text


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gridfast

    with pytest.raises(gridfast.SgkError):
        gridfast.BatchedEnv("BoatRace-v0", 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "safe-grid-agents_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libcgrid" not in src and "cg_rollout" not in src, f
