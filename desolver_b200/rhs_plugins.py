"""Built-in right-hand sides: plain Python callables TAGGED with a fused device plugin.

Each function is an ordinary `rhs(t, y, **constants)` callable in the reference's RHS protocol
(/root/reference/desolver/differential_system.py:338-344, 536-539): it works on numpy arrays and
torch tensors alike, so the unmodified reference can integrate it (that is how the fixtures in
tests/golden/ were produced).  The `device_plugin` attribute names the fused sm_100a
implementation in desolver_b200/csrc/ that `OdeSystem(..., ensemble=True)` dispatches to, and
`plugin_params` lists, in ABI order, the `constants` keys that the kernel reads per trajectory.
`DiffRHS` forwards unknown attribute reads to the wrapped callable (reference :438-439), so the
tag survives wrapping.
"""
import weakref


# ids shared with include/desolver_b200.h (DSB_RHS_*) and oracle/desolver_oracle.h (ORC_RHS_*)
RHS_LINEAR, RHS_LORENZ, RHS_VDP, RHS_NBODY, RHS_SHO, RHS_FORCED, RHS_DUFFING = 0, 1, 2, 3, 4, 5, 6


def _stack(parts, like):
    if type(like).__module__.split(".")[0] == "torch":
        import torch
        return torch.stack(parts)
    import numpy
    return numpy.stack(parts)


def lorenz(t, y, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
    """Lorenz-63.  y[0], y[1], y[2] = x, y, z (trailing axes broadcast over an ensemble)."""
    x, v, z = y[0], y[1], y[2]
    return _stack([sigma * (v - x), x * (rho - z) - v, x * v - beta * z], y)


lorenz.device_plugin = "lorenz"
lorenz.plugin_dim = (3,)          # state shape of ONE trajectory: lets OdeSystem check shapes without a probe call
lorenz.plugin_id = RHS_LORENZ
lorenz.plugin_params = ("sigma", "rho", "beta")
lorenz.plugin_defaults = dict(sigma=10.0, rho=28.0, beta=8.0 / 3.0)


def van_der_pol(t, y, mu=1.0):
    """Van der Pol oscillator, y = (x, v): x' = v, v' = mu (1 - x^2) v - x."""
    x, v = y[0], y[1]
    return _stack([v, mu * (1 - x * x) * v - x], y)


van_der_pol.device_plugin = "van_der_pol"
van_der_pol.plugin_dim = (2,)          # state shape of ONE trajectory: lets OdeSystem check shapes without a probe call
van_der_pol.plugin_id = RHS_VDP
van_der_pol.plugin_params = ("mu",)
van_der_pol.plugin_defaults = dict(mu=1.0)


def harmonic_oscillator(t, y, k=1.0, m=1.0):
    """README oscillator, y = (x, v): x' = v, v' = -(k/m) x.

    Equal, bit for bit on finite inputs, to the reference README's `[[0,1],[-k/m,0]] @ y`
    (/root/reference/docs/examples/getting_started.py:15-24): the products with 0 and 1 are exact.
    """
    return _stack([y[1], -(k / m) * y[0]], y)


harmonic_oscillator.device_plugin = "harmonic_oscillator"
harmonic_oscillator.plugin_dim = (2,)          # state shape of ONE trajectory: lets OdeSystem check shapes without a probe call
harmonic_oscillator.plugin_id = RHS_SHO
harmonic_oscillator.plugin_params = ("k", "m")
harmonic_oscillator.plugin_defaults = dict(k=1.0, m=1.0)


def linear(t, y, A=None, out=None):
    """Dense linear system y' = A @ y (A shared by the whole ensemble; y is (n,) or (n, N)).

    For an ensemble of CUDA float64 columns whose shape fits its tiling, the product is one of the library's own
    tensor-core kernels: `linear.gemm = "dmma"` the FP64 DMMA kernel (dsb_dgemm, csrc/dgemm.cu; the rounding sequence of
    an FP64 GEMM, bit-identical to cuBLAS), `linear.gemm = "int8"` the error-free digit-plane product on the INT8 tensor
    cores (dsb_ozaki_*, csrc/ozaki_gemm.cu; closer to the exact product than an FP64 GEMM and 2.4x faster than cuBLAS at
    4096 x 8192 x 4096; shapes outside its 128 x 64 x 128 tiling are zero-padded in staging buffers), `"auto"` (default) the INT8 kernel for large products (n >= 1024 and >= 512
    columns), the DMMA kernel otherwise.  Numpy arrays, the reference, odd shapes: the array library's matmul;
    `linear.use_library_gemm = True` forces the latter."""
    if (not linear.use_library_gemm and type(y).__module__.split(".")[0] == "torch" and y.is_cuda and y.ndim == 2
            and getattr(A, "is_cuda", False)):
        from .backend import cuda_backend as cb
        if y.dtype == A.dtype and _linear_wants_int8(A, y.shape[1]):
            prod = _linear_int8_product(A)
            if prod.supports(y):
                res = prod(y, out=out if (out is not None and out.is_contiguous()) else None)
                linear.calls_int8 += 1
                return res
        res = cb.dgemm(A, y, out=out if (out is not None and out.is_contiguous()) else None)
        if res is not None:
            linear.calls_dsb += 1
            return res
    linear.calls_library += 1
    if out is not None and type(y).__module__.split(".")[0] == "torch" and out.is_contiguous():
        import torch
        return torch.matmul(A, y, out=out)
    return A @ y


def _linear_wants_int8(A, columns):
    import torch
    if not (getattr(A, "is_cuda", False) and A.ndim == 2 and A.dtype == torch.float64 and A.is_contiguous()):
        return False
    return linear.gemm == "int8" or (linear.gemm == "auto" and min(A.shape) >= 1024 and columns >= 512)


def _linear_int8_product(A):
    from .backend import cuda_backend as cb
    prod = linear._int8.get(id(A))
    if prod is None or prod.A is not A:
        linear._int8.clear()                                # one matrix at a time: the planes of A are as large as A
        prod = linear._int8[id(A)] = cb.Int8Product(A)
        weakref.finalize(A, linear._int8.pop, id(A), None)  # the planes go when the matrix goes
    return prod


def _linear_prepare(A=None, **unused):
    """Called by OdeSystem at the start of every stage-level integrate(): the digit planes of an A that was modified in
    place are cut again BEFORE captured attempts are replayed (a CUDA graph replays kernels, not this Python)."""
    if not linear.use_library_gemm and linear.gemm in ("auto", "int8") and _linear_wants_int8(A, 1 << 30):
        prod = _linear_int8_product(A)
        if prod.shape_ok():
            prod.refresh()


linear.prepare = _linear_prepare
linear.use_library_gemm = False
linear.gemm = "auto"              # "auto" | "dmma" | "int8": which tensor-core kernel serves the product (see the docstring)
linear._int8 = {}
linear.calls_int8 = 0
linear.calls_dsb = 0              # products served by dsb_dgemm / by the array library (CUDA-graph replays are not
linear.calls_library = 0          # re-counted: what matters is which kernel a captured attempt contains)
linear.supports_out = True        # rhs(t, y, out=array, **constants) writes the derivative into a caller-owned array
linear.column_multiple = 128      # dsb_dgemm tiles 128 ensemble columns: compacted working sets are padded to that


linear.device_plugin = "linear"
linear.plugin_id = RHS_LINEAR
linear.plugin_params = ()
linear.plugin_shared = ("A",)


def nbody(t, y, m=None, eps2=0.0, G=1.0):
    """Softened gravitational N-body.  y = (2, nb, 3[, N]): y[0] positions, y[1] velocities; a trailing
    axis enumerates independent systems (the ensemble axis of `OdeSystem(..., ensemble=True)`).
    Masses m (nb,) are shared by all systems."""
    q, v = y[0], y[1]                                       # (nb, 3, ...)
    extra = (1,) * (q.ndim - 2)
    d = q[None] - q[:, None]                                # d[i, j] = q[j] - q[i]   (nb, nb, 3, ...)
    r2 = (d * d).sum(2) + eps2                              # (nb, nb, ...)
    nb = q.shape[0]
    if type(y).__module__.split(".")[0] == "torch":
        import torch
        off = 1.0 - torch.eye(nb, dtype=q.dtype, device=q.device)
        mm = torch.as_tensor(m, dtype=q.dtype, device=q.device)
    else:
        import numpy
        off = 1.0 - numpy.eye(nb)
        mm = numpy.asarray(m)
    w = mm.reshape((1, nb) + extra) * r2 ** -1.5
    w = w * off.reshape((nb, nb) + extra)                   # no self-interaction
    acc = G * (d * w[:, :, None]).sum(1)                    # sum over j   (nb, 3, ...)
    return _stack([v, acc], y)


nbody.device_plugin = "nbody"
nbody.plugin_id = RHS_NBODY
nbody.plugin_params = ()
nbody.plugin_shared = ("G", "eps2", "m")

def forced_oscillator(t, y):
    """y'' + y = t written as a first-order system (the reference's own test fixture,
    /root/reference/desolver/tests/common.py:27-72).  NOT tagged: it is time dependent and has no fused
    kernel, so it exercises the stage-level path (user callable between fused stage kernels)."""
    return _stack([y[1], -y[0] + t], y)


def forced(t, y):
    """The same system as `forced_oscillator`, TAGGED: time-dependent right-hand sides run in the fused kernel too
    (the stage time t + c_s h is formed per lane, /root/reference/desolver/integrators/components/
    runge_kutta_methods.py:49-50)."""
    return _stack([y[1], -y[0] + t], y)


forced.device_plugin = "forced"
forced.plugin_dim = (2,)          # state shape of ONE trajectory: lets OdeSystem check shapes without a probe call
forced.plugin_id = RHS_FORCED
forced.plugin_params = ()
forced.plugin_defaults = dict()


def _cos(x):
    if type(x).__module__.split(".")[0] == "torch":
        import torch
        return torch.cos(x)
    import numpy
    return numpy.cos(x)


def duffing(t, y, delta=0.2, alpha=-1.0, beta=1.0, gamma=0.3, omega=1.2):
    """Duffing oscillator x'' + delta x' + alpha x + beta x^3 = gamma cos(omega t), y = (x, v) -- the system of the
    reference's example notebooks (/root/reference/docs/examples).  Any parameter may be per-trajectory."""
    x, v = y[0], y[1]
    return _stack([v, gamma * _cos(omega * t) - delta * v - alpha * x - beta * x * x * x], y)


duffing.device_plugin = "duffing"
duffing.plugin_dim = (2,)          # state shape of ONE trajectory: lets OdeSystem check shapes without a probe call
duffing.plugin_id = RHS_DUFFING
duffing.plugin_params = ("delta", "alpha", "beta", "gamma", "omega")
duffing.plugin_defaults = dict(delta=0.2, alpha=-1.0, beta=1.0, gamma=0.3, omega=1.2)

# right-hand sides with a thread-per-trajectory fused kernel (csrc/rhs_builtin.cuh)
FUSED = ("lorenz", "van_der_pol", "harmonic_oscillator", "forced", "duffing")

BUILTIN = {f.device_plugin: f for f in (lorenz, van_der_pol, harmonic_oscillator, linear, nbody, forced, duffing)}
BUILTIN["forced_oscillator"] = forced_oscillator


# ---------------------------------------------------------------------------------------------------------------------
# Right-hand sides of the CALLER as fused device plugins.
#
# The reference integrates any Python callable (differential_system.py:338-344); here an untagged callable runs between
# the stage-level kernels (one CUDA graph per lock-step round: 158 ms for the 2^20-trajectory Lorenz ensemble).  A callable
# whose arithmetic the caller ALSO writes down as a CUDA expression list gets the persistent thread-per-trajectory kernel
# instead (3.5 ms for the same ensemble): the body is compiled with the kernel templates of this library
# (csrc/erk_fused.cuh, csrc/plan.h) into a plug-in shared object of its own, registered with the library under a new
# right-hand-side id (dsb_rhs_register), and the callable is tagged like the built-in ones.
_USER_PLUGINS = {}          # source hash -> (rhs id, ctypes handle of the plug-in)
_USER_BASE = 1000           # DSB_RHS_USER_BASE

_PLUGIN_SOURCE = """// generated by desolver_b200.rhs_plugins.register_device_rhs -- %(name)s
#include <cstdarg>
#include <cstdio>
#define DSB_SELECT_FAST_ONLY 1
#include "plan.h"
namespace dsb {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\\n', stderr); }
template <bool STRICT>
struct UserRhs {
    static constexpr int ID = %(base)d, DIM = %(dim)d, NPARAM = %(nparam)d;
    static constexpr bool TIME = %(time)s;
    static __device__ __forceinline__ void prepare(double (&)[NPARAM > 0 ? NPARAM : 1]) {}
    static __device__ __forceinline__ void eval(%(targ)sconst double (&y)[DIM], const double (&p)[NPARAM > 0 ? NPARAM : 1],
                                                double (&dy)[DIM])
    {
%(body)s
    }
};
}  // namespace dsb
extern "C" __attribute__((visibility("default"))) bool dsb_user_rhs_select(dsb_erk_plan *p)
{
    return dsb::select_fused<dsb::UserRhs>(p);
}
"""


class _Sym(object):
    """One scalar of the right-hand side while `trace_body` evaluates the caller's function: every operator returns the
    CUDA expression of its result, in the order Python evaluates it (so the statement list repeats the callable's
    arithmetic operation for operation)."""
    __slots__ = ("code",)
    __array_priority__ = 1000
    _FUNCS = ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "atan", "asin", "acos", "abs", "exp2", "log2",
              "log10", "log1p", "expm1", "atan2", "pow", "fmin", "fmax")
    _ALIASES = dict(arctan="atan", arcsin="asin", arccos="acos", arctan2="atan2", absolute="abs", fabs="abs", power="pow",
                    minimum="fmin", maximum="fmax", add="+", subtract="-", multiply="*", divide="/", true_divide="/", mul="*",
                    sub="-", div="/", neg="neg", negative="neg", square="square", reciprocal="reciprocal")

    def __init__(self, code):
        self.code = code

    @staticmethod
    def lit(v):
        if isinstance(v, _Sym):
            return v.code
        if hasattr(v, "item") and getattr(v, "ndim", 1) == 0:
            v = v.item()
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise TypeError("trace_body: cannot take %r into a device expression (numbers and state / parameter entries only)"
                            % (type(v).__name__,))
        return "(%s)" % repr(float(v)) if v < 0 else repr(float(v))

    def _bin(self, op, other, swap=False):
        a, b = (self.lit(other), self.code) if swap else (self.code, self.lit(other))
        return _Sym("(%s %s %s)" % (a, op, b))

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __neg__(self): return _Sym("(-%s)" % self.code)
    def __pos__(self): return self
    def __abs__(self): return _Sym("fabs(%s)" % self.code)

    def __pow__(self, o):
        if isinstance(o, (int, float)) and float(o) == int(o) and 1 <= int(o) <= 4:       # x**2 = x*x like numpy / torch
            return _Sym("(" + " * ".join([self.code] * int(o)) + ")")
        return _Sym("pow(%s, %s)" % (self.code, self.lit(o)))

    def __rpow__(self, o):
        return _Sym("pow(%s, %s)" % (self.lit(o), self.code))

    @classmethod
    def call(cls, name, args):
        name = cls._ALIASES.get(name, name)
        a = [cls.lit(v) for v in args]
        if name in ("+", "-", "*", "/") and len(a) == 2:
            return _Sym("(%s %s %s)" % (a[0], name, a[1]))
        if name == "neg":
            return _Sym("(-%s)" % a[0])
        if name == "square":
            return _Sym("(%s * %s)" % (a[0], a[0]))
        if name == "reciprocal":
            return _Sym("(1.0 / %s)" % a[0])
        if name == "abs":
            return _Sym("fabs(%s)" % a[0])
        if name in cls._FUNCS:
            return _Sym("%s(%s)" % (name, ", ".join(a)))
        raise NotImplementedError("trace_body: no device counterpart for %r" % name)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        return self.call(ufunc.__name__, inputs)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, "__name__", str(func))
        if name in ("stack", "cat", "concatenate") and not kwargs:
            return list(args[0])
        if name in ("as_tensor", "tensor", "asarray"):
            return args[0]
        return cls.call(name, args)


def trace_body(fn, dim, params=(), time_dependent=None):
    """The CUDA statement list of `fn(t, y, **constants)`: the callable is evaluated ONCE on symbolic scalars (`y[i]`, `p[k]`,
    `t`), whose operators record the expression they compute.  Works for callables made of arithmetic operators, `**` with
    small integer exponents, and elementwise numpy / torch functions with a CUDA namesake (sin, cos, exp, log, sqrt, tanh,
    ...), finished by `stack([...])` (numpy or torch) or a list -- i.e. no data-dependent control flow, no reductions.
    Returns (body, time_dependent)."""
    y = [_Sym("y[%d]" % i) for i in range(int(dim))]
    consts = {name: _Sym("p[%d]" % k) for k, name in enumerate(params)}
    out = fn(_Sym("t"), y, **consts)
    out = list(out.tolist() if hasattr(out, "tolist") else out)
    if len(out) != int(dim):
        raise ValueError("trace_body: the callable returned %d components for a state of %d" % (len(out), int(dim)))
    lines = ["dy[%d] = %s;" % (i, _Sym.lit(v)) for i, v in enumerate(out)]
    uses_t = any(_uses_time(_Sym.lit(v)) for v in out)
    return "\n".join(lines), (uses_t if time_dependent is None else bool(time_dependent))


def _uses_time(code):
    import re
    return re.search(r"(?<![A-Za-z0-9_\]])t(?![A-Za-z0-9_(\[])", code) is not None


def register_device_rhs(fn, body=None, dim=None, params=(), defaults=None, time_dependent=None, name=None, build_dir=None):
    """Tags the callable `fn(t, y, **constants)` with a fused device plugin compiled from `body`.

    `body`: CUDA C++ statements that fill `dy[0..dim)` from `y[0..dim)`, the per-trajectory parameters `p[0..len(params))`
    (the `constants` entries named in `params`, in that order; a (N,) tensor is per trajectory, a scalar is shared) and, if
    `time_dependent`, the stage time `t`; `body=None` writes them by evaluating `fn` once on symbolic scalars
    (`trace_body`: arithmetic and elementwise functions only) -- the same arithmetic as `fn`, which stays the definition of the system: the
    stage-level kernels (per-component tolerances, events with dense output, STRICT arithmetic, schemes without a fused
    instantiation) and the reference itself keep calling it.  The plug-in is built once with nvcc (about a minute: every
    tableau / event / dense-output / multi-stop instantiation of the kernel) into `build_dir` (default: `_plugins/` next
    to this file, so a prebuilt one travels with the tree) and reused by its source hash.  Returns `fn`."""
    import ctypes
    import hashlib
    import os
    import shutil
    import subprocess
    from .backend import cuda_backend as cb
    global FUSED
    if dim is None:
        raise ValueError("register_device_rhs: dim (components of ONE trajectory's state) is required")
    dim, params = int(dim), tuple(params)
    if body is None:
        body, time_dependent = trace_body(fn, dim, params, time_dependent)
    time_dependent = bool(time_dependent)
    if dim < 1 or len(params) > 6:
        raise ValueError("register_device_rhs: dim >= 1 and at most 6 per-trajectory parameters (DSB_MAX_PARAMS)")
    name = name or getattr(fn, "__name__", "rhs")
    if not name.isidentifier():
        raise ValueError("register_device_rhs: name must be an identifier")
    pkg = os.path.dirname(os.path.abspath(__file__))
    src = _PLUGIN_SOURCE % dict(name=name, base=_USER_BASE, dim=dim, nparam=len(params), time="true" if time_dependent else "false",
                                targ="double t, " if time_dependent else "",
                                body="\n".join("        " + line.strip() for line in body.strip().splitlines()))
    headers = b"".join(open(os.path.join(pkg, "csrc", h), "rb").read() for h in
                       ("plan.h", "erk_fused.cuh", "common.cuh", "fastmath.cuh", "factor_table.h", "tableau_static.cuh"))
    key = hashlib.sha256(src.encode() + headers).hexdigest()[:16]
    if key not in _USER_PLUGINS:
        out_dir = build_dir or os.path.join(pkg, "_plugins")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "dsb_rhs_%s_%s.so" % (name, key))
        if not os.path.exists(so):
            nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
            cu = os.path.join(out_dir, "dsb_rhs_%s_%s.cu" % (name, key))
            with open(cu, "w") as f:
                f.write(src)
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-diag-suppress", "128",
                   "-I", os.path.join(pkg, "csrc"), "-I", os.path.join(os.path.dirname(pkg), "include"), "-shared",
                   "-o", so + ".tmp", cu, "-lcudart"]
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            if res.returncode != 0:
                raise RuntimeError("register_device_rhs: nvcc failed for %s:\n%s" % (cu, res.stdout.decode()[-4000:]))
            os.replace(so + ".tmp", so)
        plug = ctypes.CDLL(so)
        rhs_id = _USER_BASE + len(_USER_PLUGINS)
        L = cb.lib()
        L.dsb_rhs_register.restype = ctypes.c_int
        L.dsb_rhs_register.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        cb.check(L.dsb_rhs_register(rhs_id, dim, ctypes.cast(plug.dsb_user_rhs_select, ctypes.c_void_p)), "dsb_rhs_register")
        _USER_PLUGINS[key] = (rhs_id, plug)
    rhs_id = _USER_PLUGINS[key][0]
    tag = "user:%s:%s" % (name, key)
    fn.device_plugin, fn.plugin_id, fn.plugin_params = tag, rhs_id, params
    fn.plugin_defaults = dict(defaults or {})
    fn.plugin_dim = (dim,)
    if tag not in FUSED:
        FUSED = FUSED + (tag,)
    return fn
