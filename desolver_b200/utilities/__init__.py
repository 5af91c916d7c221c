"""`desolver_b200.utilities` -- the pieces of the reference's `desolver.utilities` that the hot path's protocol exposes
(/root/reference/desolver/utilities/__init__.py): the cubic Hermite interpolant returned by `integrator.dense_output()`,
and the row lookups behind `a[t]` / `sol(t)` (`search_bisection`, `search_bisection_vec`: utilities/utilities.py:223-300).
The Jacobian wrapper and the root finders of the implicit schemes are out of scope."""
import warnings

import numpy as np

from . import interpolation
from .interpolation import CubicHermiteInterp

__all__ = ["interpolation", "CubicHermiteInterp", "search_bisection", "search_bisection_vec", "warning"]


def warning(message, category=Warning):
    """`warnings.warn` under the reference's name (utilities/utilities.py:203-220)."""
    warnings.warn(message, category=category)


def _host(array):
    if hasattr(array, "detach"):
        array = array.detach().cpu().numpy()
    return np.asarray(array)


def search_bisection_vec(array, val):
    """For every entry of `val`: the index of the first element of the ascending `array` that is >= it, clamped to the
    array (what the reference's bisection returns, utilities/utilities.py:271-300).  Device tensors are looked up on the
    host copy; the result is an int64 numpy array (or a tensor on `val`'s device if `val` is one)."""
    arr, v = _host(array), _host(val)
    idx = np.minimum(np.searchsorted(arr, v, side="left"), len(arr) - 1).astype(np.int64)
    if hasattr(val, "detach"):
        import torch
        return torch.as_tensor(idx, device=val.device)
    return idx


def search_bisection(array, val):
    """Scalar form (utilities/utilities.py:223-267): first index with array[i] >= val, clamped."""
    return int(search_bisection_vec(array, np.asarray(float(val)).reshape(1))[0])
