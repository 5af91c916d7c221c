"""`desolver_b200.backend` -- the CUDA backend namespace (`D` in the reference's vocabulary).

Mirrors the helpers of /root/reference/desolver/backend/autoray_backend.py:8-39 that the hot
path uses (`epsilon` = 4 eps, `tol_epsilon` = 32 eps, `backend_like_dtype`) and owns the loader
of the C-ABI library (`cuda_backend`).  There is no numpy/CPU arithmetic path here: arrays are
torch CUDA tensors and every step of the hot path runs in libdesolver_b200.so.
"""
from functools import lru_cache

import numpy as np

from . import cuda_backend
from .cuda_backend import lib, library_path, is_available  # noqa: F401

e = 2.718281828459045
pi = 3.141592653589793
euler_gamma = 0.5772156649015329


def _finfo_eps(dtype):
    if "torch" in str(dtype):
        import torch
        return torch.finfo(dtype).eps
    return float(np.finfo(np.dtype(dtype)).eps)


@lru_cache(maxsize=32)
def epsilon(dtype):
    """4 * machine epsilon of `dtype` (reference backend/autoray_backend.py:8-18)."""
    return _finfo_eps(dtype) * 4


@lru_cache(maxsize=32)
def tol_epsilon(dtype):
    """32 * machine epsilon of `dtype` (reference backend/autoray_backend.py:21-31)."""
    return _finfo_eps(dtype) * 32


def backend_like_dtype(dtype):
    """'torch' for torch dtypes, 'numpy' for numpy dtypes (reference :34-39)."""
    if "torch" in str(dtype):
        return "torch"
    return "numpy"
