"""ctypes front-end of the CPU oracle (oracle/desolver_oracle.c).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
`--impl reference` legs.  Nothing under desolver_b200/ imports this module; the product
path has no CPU fallback.

The C source restates the reference algorithm (file:line citations are in
desolver_oracle.h); it is pinned against the unmodified reference by the fixtures that
oracle/make_golden.py wrote into tests/golden/.
"""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libdesolver_oracle.so")

RHS_LINEAR, RHS_LORENZ, RHS_VDP, RHS_NBODY, RHS_SHO, RHS_FORCED, RHS_DUFFING = 0, 1, 2, 3, 4, 5, 6
STATUS_OK, STATUS_TOL_FAIL, STATUS_RUNNING = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class _Problem(C.Structure):
    _fields_ = [("rhs_id", C.c_int), ("dim", C.c_int), ("n_params", C.c_int),
                ("shared", _dp), ("n_shared", C.c_int), ("drift_count", C.c_int)]


class _Tableau(C.Structure):
    _fields_ = [("stages", C.c_int), ("inter", _dp), ("final_", _dp),
                ("final_rows", C.c_int), ("order", C.c_double)]


class _Dense(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("count", C.c_int64), ("t0", _dp), ("t1", _dp),
                ("y0", _dp), ("y1", _dp), ("f0", _dp), ("f1", _dp)]


_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "desolver_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        lib.orc_erk_trajectory.restype = C.c_int
        lib.orc_erk_trajectory.argtypes = [C.POINTER(_Problem), C.POINTER(_Tableau), _dp, _dp, _dp, _dp,
                                           C.c_double, C.c_double, C.c_double, C.c_int64, _lp, _lp,
                                           C.POINTER(_Dense)]
        lib.orc_symplectic_trajectory.restype = C.c_int
        lib.orc_symplectic_trajectory.argtypes = [C.POINTER(_Problem), C.c_int, _dp, _dp, _dp, _dp, _dp,
                                                  C.c_double, C.c_int64, _lp]
        lib.orc_erk_ensemble.restype = None
        lib.orc_erk_ensemble.argtypes = [C.POINTER(_Problem), C.POINTER(_Tableau), _dp, _dp, _dp, _dp,
                                         C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                         C.c_double, _ip, _ip, _ip]
        lib.orc_symplectic_ensemble.restype = None
        lib.orc_symplectic_ensemble.argtypes = [C.POINTER(_Problem), C.c_int, _dp, _dp, _dp, _dp, _dp,
                                                C.c_int64, C.c_int64, C.c_int64, C.c_double, _ip, _ip]
        lib.orc_hermite_eval.restype = None
        lib.orc_hermite_eval.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, C.c_double, _dp]
        lib.orc_dense_eval.restype = None
        lib.orc_dense_eval.argtypes = [C.POINTER(_Dense), C.c_int, C.c_double, _dp]
        lib.orc_set_factor_perturbation.restype = None
        lib.orc_set_factor_perturbation.argtypes = [C.c_int]
        lib.orc_rhs.restype = None
        lib.orc_rhs.argtypes = [C.POINTER(_Problem), _dp, C.c_double, _dp, _dp]
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _problem(rhs_id, dim, n_params, shared):
    sh = _f64(shared) if shared is not None else None
    prob = _Problem(rhs_id, dim, n_params, _ptr(sh), 0 if sh is None else sh.size, -1)
    return prob, sh


def _tableau(inter, final, order):
    inter, final = _f64(inter), _f64(np.atleast_2d(final))
    tab = _Tableau(inter.shape[0], _ptr(inter), _ptr(final), final.shape[0], float(order))
    return tab, (inter, final)


def _shards(n, threads):
    threads = max(1, min(int(threads), n))
    edges = np.linspace(0, n, threads + 1).astype(np.int64)
    return [(int(edges[i]), int(edges[i + 1] - edges[i])) for i in range(threads)]


def _run_sharded(fn, n, threads):
    shards = _shards(n, threads)
    if len(shards) == 1:
        fn(*shards[0])
        return
    ts = [threading.Thread(target=fn, args=s) for s in shards]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


def erk_ensemble(rhs_id, y0, t0, tf, dt0, rtol, atol, inter, final, order, params=None, shared=None,
                 threads=1):
    """Integrate every column of y0 (dim, N) independently with reference semantics.

    params: (P, N) per-trajectory parameters (or (P,) broadcast).  Returns a dict with the final
    SoA state `y` (dim, N), `t`, `dt`, int32 `accepted`, `rejected`, `status`.
    """
    lib = load()
    y = _f64(np.array(y0, dtype=np.float64, copy=True))
    squeeze = y.ndim == 1
    if squeeze:
        y = y[:, None].copy()
    dim, n = y.shape
    prm = np.zeros((0, n)) if params is None else np.array(params, dtype=np.float64)
    if prm.ndim == 1:
        prm = np.repeat(prm[:, None], n, axis=1)
    prm = _f64(prm)
    t = _f64(np.broadcast_to(np.float64(t0), (n,)).copy())
    dt = _f64(np.broadcast_to(np.float64(dt0), (n,)).copy())
    acc = np.zeros(n, np.int32)
    rej = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32)
    prob, _sh = _problem(rhs_id, dim, prm.shape[0], shared)
    tab, _keep = _tableau(inter, final, order)

    def work(first, count):
        lib.orc_erk_ensemble(C.byref(prob), C.byref(tab), _ptr(prm), _ptr(y), _ptr(t), _ptr(dt),
                             n, first, count, float(tf), float(rtol), float(atol),
                             acc.ctypes.data_as(_ip), rej.ctypes.data_as(_ip), st.ctypes.data_as(_ip))
    _run_sharded(work, n, threads)
    if squeeze:
        y = y[:, 0]
    return dict(y=y, t=t, dt=dt, accepted=acc, rejected=rej, status=st)


def symplectic_ensemble(rhs_id, y0, t0, tf, dt0, tab3, params=None, shared=None, threads=1, drift_count=-1):
    """Fixed-step symplectic splitting over every column of y0 (dim, N).  `drift_count` leading state elements
    are drift variables (default dim // 2; (dim0 // 2) * prod(dim[1:]) for a state whose axis 0 is odd)."""
    lib = load()
    y = _f64(np.array(y0, dtype=np.float64, copy=True))
    squeeze = y.ndim == 1
    if squeeze:
        y = y[:, None].copy()
    dim, n = y.shape
    prm = np.zeros((0, n)) if params is None else np.array(params, dtype=np.float64)
    if prm.ndim == 1:
        prm = np.repeat(prm[:, None], n, axis=1)
    prm = _f64(prm)
    t = _f64(np.broadcast_to(np.float64(t0), (n,)).copy())
    dt = _f64(np.broadcast_to(np.float64(dt0), (n,)).copy())
    steps = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32)
    prob, _sh = _problem(rhs_id, dim, prm.shape[0], shared)
    prob.drift_count = int(drift_count)
    tab3 = _f64(tab3)

    def work(first, count):
        lib.orc_symplectic_ensemble(C.byref(prob), tab3.shape[0], _ptr(tab3), _ptr(prm), _ptr(y), _ptr(t),
                                    _ptr(dt), n, first, count, float(tf),
                                    steps.ctypes.data_as(_ip), st.ctypes.data_as(_ip))
    _run_sharded(work, n, threads)
    if squeeze:
        y = y[:, 0]
    return dict(y=y, t=t, dt=dt, steps=steps, status=st)


class DenseTrajectory:
    """Recorded (t0, t1, y0, y1, f0, f1) nodes of ONE trajectory plus Hermite evaluation."""

    def __init__(self, dim, capacity):
        self.dim = dim
        self.t0 = np.zeros(capacity)
        self.t1 = np.zeros(capacity)
        self.y0 = np.zeros((capacity, dim))
        self.y1 = np.zeros((capacity, dim))
        self.f0 = np.zeros((capacity, dim))
        self.f1 = np.zeros((capacity, dim))
        self.c = _Dense(capacity, 0, _ptr(self.t0), _ptr(self.t1), _ptr(self.y0), _ptr(self.y1),
                        _ptr(self.f0), _ptr(self.f1))

    @property
    def count(self):
        return int(self.c.count)

    def __call__(self, tq):
        out = np.zeros(self.dim)
        load().orc_dense_eval(C.byref(self.c), self.dim, float(tq), _ptr(out))
        return out


def erk_trajectory(rhs_id, y0, t0, tf, dt0, rtol, atol, inter, final, order, params=None, shared=None,
                   max_steps=-1, dense_capacity=0):
    """One trajectory; optionally records dense-output nodes.  Returns a dict."""
    lib = load()
    y = _f64(np.array(y0, dtype=np.float64, copy=True)).reshape(-1)
    prm = _f64(np.zeros(1) if params is None else np.array(params, dtype=np.float64).reshape(-1))
    n_params = 0 if params is None else prm.size
    prob, _sh = _problem(rhs_id, y.size, n_params, shared)
    tab, _keep = _tableau(inter, final, order)
    t = C.c_double(float(t0))
    dt = C.c_double(float(dt0))
    acc = C.c_int64(0)
    att = C.c_int64(0)
    dense = DenseTrajectory(y.size, dense_capacity) if dense_capacity > 0 else None
    st = lib.orc_erk_trajectory(C.byref(prob), C.byref(tab), _ptr(prm), _ptr(y), C.byref(t), C.byref(dt),
                                float(tf), float(rtol), float(atol), int(max_steps), C.byref(acc),
                                C.byref(att), C.byref(dense.c) if dense is not None else None)
    return dict(y=y, t=t.value, dt=dt.value, accepted=acc.value, attempts=att.value, status=st, dense=dense)


def rhs(rhs_id, y, params=None, shared=None, t=0.0):
    lib = load()
    y = _f64(y).reshape(-1)
    prm = _f64(np.zeros(1) if params is None else np.array(params, dtype=np.float64).reshape(-1))
    prob, _sh = _problem(rhs_id, y.size, 0 if params is None else prm.size, shared)
    out = np.zeros_like(y)
    lib.orc_rhs(C.byref(prob), _ptr(prm), float(t), _ptr(y), _ptr(out))
    return out


def hermite(t0, t1, p0, p1, m0, m1, tq):
    p0, p1, m0, m1 = (_f64(a).reshape(-1) for a in (p0, p1, m0, m1))
    out = np.zeros_like(p0)
    load().orc_hermite_eval(p0.size, float(t0), float(t1), _ptr(p0), _ptr(p1), _ptr(m0), _ptr(m1),
                            float(tq), _ptr(out))
    return out


def set_factor_perturbation(ulps):
    """Test knob of the C oracle: move every step-size factor by `ulps` units in the last place (0 = off)."""
    load().orc_set_factor_perturbation(int(ulps))
