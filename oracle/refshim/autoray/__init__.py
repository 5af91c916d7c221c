"""Dispatch-only stand-in for the `autoray` package (TEST INFRASTRUCTURE, not product code).

The reference (Microno95/desolver v5.1.0) imports `autoray` at
desolver/backend/autoray_backend.py:1, and `autoray` is not installed in the build
container (no network).  This module provides just the lazy "call the numpy-API
function of whichever library owns the first argument" dispatch that the reference
uses, so that the UNMODIFIED reference under /root/reference can be imported by
`oracle/make_golden.py` to generate the fixtures in tests/golden/.

It performs no arithmetic of its own: every numeric result is produced by numpy
(or torch) exactly as the reference calls them.  Only `oracle/make_golden.py`
puts this directory on sys.path; nothing under desolver_b200/ imports it.
"""
import contextlib
import sys
import types

import numpy as _np

from . import autoray as autoray  # noqa: F401  (holds _FUNC_ALIASES like the real package)

__version__ = "0.0+dispatch-standin"

_registered = {}
_backend_stack = []


def infer_backend(x):
    if isinstance(x, str):
        return x
    root = type(x).__module__.split(".")[0]
    return root if root in ("numpy", "torch") else "builtins"


def _pick_backend(args, kwargs):
    like = kwargs.pop("like", None)
    if like is not None:
        return infer_backend(like)
    if _backend_stack:
        return _backend_stack[-1]
    if args:
        first = args[0]
        if isinstance(first, (list, tuple)) and first and infer_backend(first[0]) != "builtins":
            return infer_backend(first[0])
        return infer_backend(first)
    return "numpy"


@contextlib.contextmanager
def backend_like(x):
    _backend_stack.append(infer_backend(x))
    try:
        yield
    finally:
        _backend_stack.pop()


def to_backend_dtype(name, like=None):
    if not isinstance(name, str):
        return name
    if like is not None and infer_backend(like) == "torch":
        import torch
        return getattr(torch, name)
    return getattr(_np, name)


def register_function(backend, name, fn, wrap=False):
    if wrap:
        fn = fn(_lookup(backend, name))
    _registered[(backend, name)] = fn


def _torch_translation(name):
    import torch

    def axis_to_dim(f):
        def g(*a, **k):
            if "axis" in k:
                k["dim"] = k.pop("axis")
            return f(*a, **k)
        return g

    def fix_dtype(k):
        if isinstance(k.get("dtype"), str):
            k["dtype"] = getattr(torch, k["dtype"])
        return k

    table = {
        "asarray": lambda x, **k: torch.as_tensor(x, **fix_dtype(k)),
        "array": lambda x, **k: torch.tensor(x, **fix_dtype(k)),
        "copy": lambda x: x.clone(),
        "clone": lambda x: x.clone(),
        "concatenate": axis_to_dim(torch.cat),
        "stack": axis_to_dim(torch.stack),
        "sum": axis_to_dim(torch.sum),
        "any": axis_to_dim(torch.any),
        "all": axis_to_dim(torch.all),
        "max": lambda x, **k: torch.max(x) if not k else torch.amax(x, dim=k["axis"]),
        "min": lambda x, **k: torch.min(x) if not k else torch.amin(x, dim=k["axis"]),
        "astype": lambda x, dt, **k: x.to(to_backend_dtype(dt, like="torch")),
        "to_numpy": lambda x: x.numpy(),
        "shape": lambda x: tuple(x.shape),
        "zeros": lambda s, **k: torch.zeros(s, **fix_dtype(k)),
        "ones": lambda s, **k: torch.ones(s, **fix_dtype(k)),
        "zeros_like": lambda x, **k: torch.zeros_like(x, **fix_dtype(k)),
        "ones_like": lambda x, **k: torch.ones_like(x, **fix_dtype(k)),
        "eye": lambda n, **k: torch.eye(n, **fix_dtype(k)),
        "arange": lambda *a, **k: torch.arange(*a, **fix_dtype(k)),
        "clip": lambda x, min=None, max=None: torch.clamp(x, min=min, max=max),
        "transpose": lambda x, axes=None: x.permute(*axes),
        "sort": lambda x: torch.sort(x)[0],
        "nonzero": lambda x: torch.nonzero(x, as_tuple=True),
        "no_grad": lambda: torch.no_grad(),
        "finfo": torch.finfo,
        "pow": torch.pow,
        "reshape": lambda x, s: torch.reshape(x, s),
    }
    return table[name] if name in table else getattr(torch, name)


def _lookup(backend, name):
    name = autoray._FUNC_ALIASES.get((backend, name), name)
    if (backend, name) in _registered:
        return _registered[(backend, name)]
    if backend == "builtins":
        if ("numpy", name) in _registered:
            return _registered[("numpy", name)]
        backend = "numpy"
    if backend == "numpy":
        special = {
            "to_numpy": lambda x: _np.asarray(x),
            "astype": lambda x, dt, **k: _np.asarray(x).astype(dt),
            "pow": _np.power,
        }
        return special[name] if name in special else getattr(_np, name)
    if backend == "torch":
        return _torch_translation(name)
    raise ValueError("unknown backend %r" % (backend,))


class _SubNamespace:
    """`autoray.numpy.linalg.<fn>` -> numpy.linalg.<fn> / torch.linalg.<fn>."""

    def __init__(self, sub):
        self._sub = sub

    def __getattr__(self, name):
        sub = self._sub

        def call(*args, **kwargs):
            if _pick_backend(args, kwargs) == "torch":
                import torch
                return getattr(getattr(torch, sub), name)(*args, **kwargs)
            return getattr(getattr(_np, sub), name)(*args, **kwargs)
        return call


class _LazyNumpy(types.ModuleType):
    linalg = _SubNamespace("linalg")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def call(*args, **kwargs):
            return _lookup(_pick_backend(args, kwargs), name)(*args, **kwargs)
        call.__name__ = name
        return call


numpy = _LazyNumpy("autoray.numpy")
sys.modules["autoray.numpy"] = numpy
