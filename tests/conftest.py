import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def assert_close_rowscale(y, y64, rowscale, rtol=1e-5, what=""):
    """SURVEY.md section 8d fp32 tolerance: |y - y64| <= rtol*|y64| + rtol*rowscale, where
    rowscale = sum_k |term_k| of the reduction behind each output element."""
    y = np.asarray(y, dtype=np.float64)
    y64 = np.asarray(y64, dtype=np.float64)
    bound = rtol * np.abs(y64) + rtol * np.asarray(rowscale, dtype=np.float64) + 1e-30
    err = np.abs(y - y64)
    worst = np.max(err / bound) if err.size else 0.0
    assert np.all(np.isfinite(y)), f"{what}: non-finite values"
    assert worst <= 1.0, f"{what}: error {worst:.3f}x the tolerance (max abs err {err.max():.3e})"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
