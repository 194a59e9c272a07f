import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "safe-grid-agents_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_files():
    names = sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("train_"))
    assert names, "golden fixtures missing"
    return [os.path.join(GOLDEN_DIR, n) for n in names]


@pytest.fixture(scope="session")
def train_golden_files():
    """Scalar logs of the reference's real train.train (tests/golden/make_train_golden.py)."""
    names = sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("train_"))
    assert names, "train golden fixtures missing"
    return [os.path.join(GOLDEN_DIR, n) for n in names]
