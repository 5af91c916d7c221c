"""`desolver_b200.utilities` -- the pieces of the reference's `desolver.utilities` that the hot path's protocol exposes
(/root/reference/desolver/utilities/__init__.py): the cubic Hermite interpolant returned by `integrator.dense_output()`."""
from . import interpolation
from .interpolation import CubicHermiteInterp

__all__ = ["interpolation", "CubicHermiteInterp"]
