import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def rel_inf(a, b, axis=0):
    """||a-b||_inf / ||b||_inf per trajectory (norm-relative, SURVEY.md section 7 'hard parts')."""
    a, b = np.asarray(a), np.asarray(b)
    num = np.max(np.abs(a - b), axis=axis)
    den = np.max(np.abs(b), axis=axis)
    return num / np.where(den == 0, 1.0, den)
