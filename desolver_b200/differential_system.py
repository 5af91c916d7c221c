"""`OdeSystem` of the drop-in: the reference's constructor and `integrate()` surface
(/root/reference/desolver/differential_system.py:457-1199) in front of the B200 ensemble engine.

What stays source-compatible: `OdeSystem(equ_rhs, y0, t=(0, 1), dense_output=False, dt=1.0,
rtol=None, atol=None, constants=dict())`, `.integrate(t=None, callback=None, eta=False,
events=None)`, `.method` / `.set_method`, `.y .t .dt .t0 .tf .rtol .atol .constants .nfev .njev
.sol .success .integration_status`, `len(a)`, `a[i]`, `a[t]`, `reset()`.

What is new (the reference has no ensemble concept): `ensemble=True` declares
the LAST axis of `y0` to enumerate independent trajectories.  Each column then evolves exactly as
the reference would evolve it alone -- own t, own dt, own error norm, own accept/reject chain --
and `constants` entries of trailing length N are per-trajectory.  `ensemble=False` (default)
keeps the reference meaning: the whole of `y0` is one system with one dt (an ensemble of N = 1).

History model: the reference stores every accepted step (`:1015-1016`), which is impossible for
2^20 trajectories with ragged step counts.  Here `a.y` / `a.t` hold one row per `integrate()`
call (row 0 = initial condition), `a[-1]` is the final ensemble state, per-step history is
available through `dense_output=True` (`a.sol(t)`), and the per-trajectory integer tensors
`accepted`, `rejected`, `status` replace what the reference leaves implicit in `len(a)`.
"""
import collections
import ctypes as C

import numpy as np

from . import backend as D
from . import exception_types as etypes
from . import integrators
from . import rhs_plugins
from .backend import cuda_backend as cb

__all__ = ["DiffRHS", "rhs_prettifier", "OdeSystem", "DenseOutput", "StateTuple", "solve_ivp"]

StateTuple = collections.namedtuple("StateTuple", ["t", "y", "event"])

# a stuck trajectory (dt underflow) loops forever in the reference; the kernels stop after this
# many accepted steps per integrate() call and report DSB_TRAJ_RUNNING instead
DEFAULT_STEP_CAP = 50_000_000


def _torch():
    import torch
    return torch


class DiffRHS(object):
    """Right-hand-side wrapper, same call protocol and counters as the reference's
    (`differential_system.py:293-447`): `rhs(t, y, **constants)`, `nfev`, attribute forwarding to
    the wrapped callable (which is how the `device_plugin` tag of rhs_plugins survives wrapping)."""

    def __init__(self, rhs, equ_repr=None, md_repr=None):
        self.__dict__["rhs"] = rhs
        self.__dict__["equ_repr"] = str(equ_repr) if equ_repr is not None else str(rhs)
        self.__dict__["md_repr"] = md_repr if md_repr is not None else self.__dict__["equ_repr"]
        self.__dict__["nfev"] = 0
        self.__dict__["njev"] = 0

    def __call__(self, t, y, *args, **kwargs):
        val = self.rhs(t, y, *args, **kwargs)
        self.__dict__["nfev"] += 1
        return val

    def __getattr__(self, name):
        return getattr(self.__dict__["rhs"], name)

    def __setattr__(self, name, val):
        if name in ("rhs", "equ_repr", "md_repr", "nfev", "njev"):
            self.__dict__[name] = val
        else:
            setattr(self.__dict__["rhs"], name, val)

    def __str__(self):
        return self.equ_repr

    def _repr_markdown_(self):
        return self.md_repr

    def __repr__(self):
        return "<DiffRHS({},{},{})>".format(repr(self.rhs), self.equ_repr, self.md_repr)


def rhs_prettifier(equ_repr=None, md_repr=None):
    def rhs_wrapper(rhs):
        return DiffRHS(rhs, equ_repr, md_repr)
    return rhs_wrapper


class DenseOutput(object):
    """Dense output over the node store (t_k, y_k, f_k) written by the step kernels (`node_store.NodeStore`).

    `sol(t)` evaluates, for every trajectory, the cubic Hermite interpolant of the accepted step that contains `t` --
    the lookup and formula of the reference's DenseOutput / CubicHermiteInterp
    (`differential_system.py:205-227`, `utilities/interpolation.py:41-61`) -- as one kernel launch for ALL query times:

      single system   `sol(0.3)` -> (*dim);  `sol(t_array)` -> t_array.shape + (*dim)   (the reference's shapes)
      ensemble        `sol(0.3)` -> (*dim, N);  `sol(t)` with t of shape (N,) -> one time per trajectory, (*dim, N);
                      `sol.grid(times)` with times of shape (T,) -> (T, *dim, N), every trajectory at every time;
                      `sol(t)` with t of shape (T, N) -> (T, *dim, N)
    `sol.grad(t)` is the interpolant's time derivative (`utilities/interpolation.py:63-78`), same shapes.
    """

    def __init__(self, system):
        self._sys = system

    def __len__(self):
        st = self._sys._node_store
        return 0 if st is None else int(st.node_count.max().item())

    @property
    def node_count(self):
        st = self._sys._node_store
        return None if st is None else st.node_count

    @property
    def t_min(self):
        return self._sys.t0

    def _store(self):
        st = self._sys._node_store
        if st is None or not st.with_f:
            raise ValueError("No interpolant has been added and time interval is not defined!")
        return st

    def _eval(self, t, derivative, grid=False):
        torch = _torch()
        s = self._sys
        st = self._store()
        tq = torch.as_tensor(t, dtype=torch.float64).to(s.device)
        n = s.n_traj
        if not s.ensemble:
            # reference shapes: scalar -> (*dim), array -> t.shape + (*dim)
            flat = tq.reshape(-1).contiguous()
            out = st.evaluate(flat, 0, 1, flat.numel(), derivative)             # (T, D, 1)
            return out.reshape(tuple(tq.shape) + tuple(s.dim))
        if grid:
            flat = tq.reshape(-1).contiguous()
            out = st.evaluate(flat, 0, 1, flat.numel(), derivative)             # (T, D, N)
            return out.reshape((flat.numel(),) + tuple(s.dim) + (n,))
        if tq.ndim == 0:
            out = st.evaluate(tq.reshape(1), 0, 0, 1, derivative)
            return s._user_shape(out[0])
        if tq.ndim == 1:
            if tq.shape[0] != n:
                raise ValueError("expected a scalar, one query time per trajectory (shape (%d,)), or a (T, %d) array; "
                                 "use sol.grid(times) for a time grid shared by all trajectories" % (n, n))
            out = st.evaluate(tq.contiguous(), 1, 0, 1, derivative)
            return s._user_shape(out[0])
        if tq.ndim == 2 and tq.shape[1] == n:
            tq = tq.contiguous()
            out = st.evaluate(tq, 1, n, tq.shape[0], derivative)
            return out.reshape((tq.shape[0],) + tuple(s.dim) + (n,))
        raise ValueError("unsupported query shape %r" % (tuple(tq.shape),))

    def __call__(self, t):
        return self._eval(t, False)

    def grad(self, t):
        return self._eval(t, True)

    def grid(self, times, derivative=False):
        """Every trajectory at every time of `times` (T,): (T, *dim, N) -- ONE launch of the T x N evaluation kernel."""
        return self._eval(times, derivative, grid=True)


def _capture_graph(device, body):
    """Record `body()` (kernel launches on the current stream) into a torch.cuda.CUDAGraph and return it.

    The steps of `torch.cuda.graph(g)` without its `torch.cuda.empty_cache()`: that call returns every cached block to the
    driver in front of EACH capture, and with the multi-gigabyte temporaries of a compacted working set in the cache it
    stalled one integrate() in four for 0.1-0.9 s (cudaFree, then cudaMalloc again); a capture on a side stream needs
    neither an empty cache nor a collected heap."""
    torch = _torch()
    g = torch.cuda.CUDAGraph()
    side = _capture_streams.get(device)
    if side is None:
        side = _capture_streams[device] = torch.cuda.Stream(device)
    current = torch.cuda.current_stream(device)
    side.wait_stream(current)
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            body()
        except BaseException:
            try:
                g.capture_end()
            except Exception:
                pass
            raise
        g.capture_end()
    current.wait_stream(side)
    return g


_capture_streams = {}


def _state_property(name):
    """Per-trajectory state tensor that is filled on first use: `reset()` only marks the state as pristine, and a
    fused run from a pristine state lets the kernel initialise everything itself (dsb_erk_run_args.fresh), so
    re-running an ensemble costs no fill / copy launches.  Any other reader triggers `_materialize()` first."""
    raw = "_raw_" + name

    def getter(self):
        self._materialize()
        return getattr(self, raw)

    def setter(self, value):
        self._materialize()
        setattr(self, raw, value)

    return property(getter, setter)


class OdeSystem(object):
    """Ordinary differential equation system integrated on a B200 by libdesolver_b200.so."""

    _y = _state_property("y")
    _t = _state_property("t")
    _dt = _state_property("dt")
    def _host_counter(name):                                  # noqa: N805  (class-body helper, like _state_property)
        """accepted / rejected: as _state_property, except that a host_io system whose caller supplied `host_out` WITHOUT
        this array does not download it with every pass -- it is fetched from the device arrays on first use."""
        raw, dev = "_raw_" + name, "_dev_" + name

        def getter(self):
            self._materialize()
            if getattr(self, raw) is None:
                setattr(self, raw, getattr(self, dev).cpu())
            return getattr(self, raw)

        def setter(self, value):
            self._materialize()
            setattr(self, raw, value)
        return property(getter, setter)

    accepted = _host_counter("accepted")
    rejected = _host_counter("rejected")
    del _host_counter

    @property
    def status(self):
        """Per-trajectory status words (cuda_backend.TRAJ_*).  A host_io system without a caller-supplied status buffer
        does not download them: after a clean run (the kernel's counters report no failed / unfinished trajectory) they
        are all TRAJ_DONE and are materialised here on demand; otherwise they are fetched from the device."""
        self._materialize()
        if self._raw_status is None:
            torch = _torch()
            if self._host_status_clean:
                self._raw_status = torch.zeros(self.n_traj, dtype=torch.int32)
            else:
                self._raw_status = self._dev_status.cpu()
        return self._raw_status

    @status.setter
    def status(self, value):
        self._materialize()
        self._raw_status = value

    def __init__(self, equ_rhs, y0, t=(0, 1), dense_output=False, dt=1.0, rtol=None, atol=None, constants=dict(),
                 ensemble=False, device=None, dense_capacity=None, strict=False, host_io=False, host_chunks=16,
                 host_out=None, history=None):
        torch = _torch()
        if len(t) != 2:
            raise ValueError("Two time bounds are required, only {} were given.".format(len(t)))
        if not callable(equ_rhs):
            raise TypeError("equ_rhs is not callable, please pass a callable object for the right hand side.")
        cb.lib()                                   # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("desolver_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.equ_rhs = equ_rhs if isinstance(equ_rhs, DiffRHS) else DiffRHS(equ_rhs)

        if isinstance(y0, torch.Tensor) and y0.is_cuda and device is None:
            device = y0.device
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        # host_io=True: y0 stays in (pinned) host memory and the results come back to host memory -- like the
        # reference, where the state lives wherever y0 lives (differential_system.py:507-515).  integrate() then
        # streams the ensemble through the device in `host_chunks` pieces: the upload of piece c+1 and the
        # download of piece c-1 overlap the kernel of piece c (see _run_host_pipeline).
        self.host_io = bool(host_io)
        self.host_chunks = int(host_chunks)
        y0 = torch.as_tensor(y0, dtype=torch.float64)
        if self.host_io:
            if y0.is_cuda or not ensemble:
                raise ValueError("host_io=True needs ensemble=True and a host (CPU, ideally pinned) y0")
        else:
            y0 = y0.to(self.device)
        self.ensemble = bool(ensemble)
        if self.ensemble:
            if y0.ndim < 2:
                raise ValueError("ensemble=True needs y0 of shape (*dim, N) with the trajectories on the last axis")
            self.dim = tuple(y0.shape[:-1])
            self.n_traj = int(y0.shape[-1])
        else:
            self.dim = tuple(y0.shape)
            self.n_traj = 1
        self.numel = int(np.prod(self.dim)) if len(self.dim) else 1
        self.dtype = torch.float64
        self.strict = bool(strict)

        self.__rtol = rtol
        self.__atol = atol
        self.__consts = self._device_constants(constants)
        self.__t0 = float(t[0])
        self.__tf = float(t[1])
        self.__dt0 = float(dt)
        self.__dense_output = bool(dense_output)
        # history="steps": `a.t`, `a.y`, `len(a)`, `a[i]` hold EVERY accepted step like the reference's
        # (differential_system.py:1015-1018, 1163-1196) -- the default for a single system; history="calls": one row
        # per integrate() call (row 0 = y0) -- the default for an ensemble, whose per-step history is ragged (use
        # dense_output=True and `a.sol` / `a.steps(i)` there)
        history = history or ("calls" if ensemble else "steps")
        if history not in ("steps", "calls"):
            raise ValueError("history must be 'steps' or 'calls'")
        if history == "steps" and ensemble:
            raise ValueError("history='steps' is the single-system history; an ensemble's per-step history is ragged: "
                             "use dense_output=True and a.sol(t) / a.steps(i)")
        self._record_steps = history == "steps" and not host_io
        self._dense_capacity = int(dense_capacity) if dense_capacity is not None else None
        self.__method = integrators.RK45CKSolver
        self.__int_status = 0
        self.__events = []
        self._step_cap = DEFAULT_STEP_CAP
        self._launches = 0
        self._scalar_cache = {}
        self.sync_every = 8          # stage-level path: host checks the active count every N attempts
        self.use_fused = True        # False forces the stage-level path even for a tagged right-hand side
        self.compaction = True       # stage-level path: gather the active trajectories when < 50 % remain
        self.use_graphs = True       # stage-level path: replay one attempt as a CUDA graph
        self._compactions = 0
        self.callback_every = 1
        self.event_capacity = 64     # event records kept per trajectory (device event detection)
        self._event_log = None
        # lazy_status=True: integrate() on the fused path returns as soon as the launch is enqueued -- no read-back of
        # the kernel's counters, hence no host synchronisation; tolerance failures are raised (and `stats`,
        # `success`, `integration_status` resolved) when one of them is first consulted
        self.lazy_status = False
        self._pending = False
        self._scatter_store = None   # set by distributed.ShardedEnsemble(scatter_store=True)
        self._push_target = None     # set by distributed.ShardedEnsemble(push=True)
        self.push_chunks = 16

        if self.host_io:
            self._init_host_io(y0, host_out)
            self._reset_state()
            self.initialise_integrator()
            return
        self._y0 = y0.reshape(self.numel, self.n_traj).contiguous().clone()
        # one RHS evaluation as a shape check, like the reference (:536-539) -- except for a built-in right-hand side of
        # known state shape, where comparing shapes needs no evaluation (for an ensemble the probe would run the Python
        # callable over all N trajectories: aten kernels and a full-size temporary per construction)
        known = getattr(self.equ_rhs, "plugin_dim", None) if self._fused_plugin_tag() is not None else None
        if known is not None:
            if tuple(known) != tuple(self.dim):
                raise ValueError("Expected rhs to return a shape that is compatible with y0, got shapes "
                                 "Shape[y0]={} and Shape[dy]={}".format(tuple(self._user_shape(self._y0).shape),
                                                                        tuple(known) + ((self.n_traj,) if self.ensemble else ())))
            self._param_tensors(tuple(self.equ_rhs.plugin_params), dict(self.equ_rhs.plugin_defaults))   # shapes of the constants
            self.equ_rhs.nfev += 1                          # the reference's constructor evaluates the rhs once
        else:
            probe = self.equ_rhs(self._user_time(torch.full((self.n_traj,), self.__t0, dtype=torch.float64,
                                                            device=self.device)),
                                 self._user_shape(self._y0), **self.__consts)
            if tuple(probe.shape) != tuple(self._user_shape(self._y0).shape):
                raise ValueError("Expected rhs to return a shape that is compatible with y0, got shapes "
                                 "Shape[y0]={} and Shape[dy]={}".format(tuple(self._user_shape(self._y0).shape),
                                                                        tuple(probe.shape)))
        self._reset_state()
        self.initialise_integrator()

    # ------------------------------------------------------------------ state helpers
    def _device_constants(self, constants):
        """Array-valued constants (numpy arrays, CPU tensors) move to the system's device once, so
        that both the Python callable and the fused kernels see device memory."""
        torch = _torch()
        out = {}
        for k, v in (constants or {}).items():
            if isinstance(v, np.ndarray) and v.dtype.kind == "f":
                v = torch.as_tensor(v, dtype=torch.float64)
            if isinstance(v, torch.Tensor):
                v = v.to(self.device)
            out[k] = v
        return out

    def _user_shape(self, flat):
        """(numel, N) storage -> the caller's (*dim[, N]) view."""
        return flat.reshape(*self.dim, self.n_traj) if self.ensemble else flat.reshape(self.dim)

    def _user_time(self, tvec):
        return tvec if self.ensemble else tvec.reshape(())

    def _init_host_io(self, y0, host_out):
        """host_io=True: host-side state (y0 as given, pinned result buffers) and device staging buffers."""
        torch = _torch()
        n, dev = self.n_traj, self.device
        self._y0 = y0.reshape(self.numel, n)                      # host view, never copied
        if not self._y0.is_contiguous():
            self._y0 = self._y0.contiguous()
        if self._fused_plugin_tag() is None:
            raise NotImplementedError("host_io=True is implemented for right-hand sides with a fused device plugin "
                                      "(rhs_plugins.FUSED)")
        host_out = host_out or {}

        def host_buffer(name, shape, dtype):
            buf = host_out.get(name)
            if buf is None:
                buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            if tuple(buf.shape) != tuple(shape) or buf.dtype != dtype or buf.is_cuda or not buf.is_contiguous():
                raise ValueError("host_out[%r] must be a contiguous host tensor of shape %r, dtype %s" % (name, shape, dtype))
            return buf

        self._raw_y = host_buffer("y", (self.numel, n), torch.float64)
        # counters: with a caller-supplied host_out only the arrays it contains are downloaded with every pass; the others
        # stay on the device and are fetched on first use (`a.accepted`, `a.rejected`, `a.status`)
        explicit = bool(host_out)
        self._host_acc_buffer = host_buffer("accepted", (n,), torch.int32) if (not explicit or "accepted" in host_out) else None
        self._host_rej_buffer = host_buffer("rejected", (n,), torch.int32) if (not explicit or "rejected" in host_out) else None
        self._raw_accepted, self._raw_rejected = self._host_acc_buffer, self._host_rej_buffer
        # status words: downloaded only into a buffer the caller supplies; otherwise on demand (see `status`)
        self._host_status_buffer = host_buffer("status", (n,), torch.int32) if "status" in host_out else None
        self._raw_status = self._host_status_buffer
        self._host_status_clean = True
        self._dev_y = torch.empty((self.numel, n), dtype=torch.float64, device=dev)
        self._raw_t = torch.empty((n,), dtype=torch.float64, device=dev)      # t and dt stay on the device
        self._raw_dt = torch.empty((n,), dtype=torch.float64, device=dev)
        self._dev_accepted = torch.empty(n, dtype=torch.int32, device=dev)
        self._dev_rejected = torch.empty(n, dtype=torch.int32, device=dev)
        self._dev_status = torch.empty(n, dtype=torch.int32, device=dev)
        # both are cleared by every run (counters.zero_() / the library's memset of the flags)
        self._counters = torch.empty(8 * max(1, self.host_chunks), dtype=torch.int64, device=dev)
        self._pipe_flags = torch.empty(2 + max(1, self.host_chunks) * 2, dtype=torch.int32, device=dev)
        self._pipe_streams = None
        self.host_pipeline = "streamed"      # "chunked": one dsb_erk_run launch per chunk (fallback without stream mem-ops)

    def _fused_plugin_tag(self):
        rhs = self.equ_rhs
        if getattr(rhs, "plugin_id", None) is None:
            return None
        if getattr(rhs, "device_plugin", None) not in rhs_plugins.FUSED:
            return None
        return rhs.plugin_id

    def _initial_dt(self):
        dt0 = self.__dt0                                  # sign follows the direction of integration (reference :728-732)
        if np.sign(dt0) != np.sign(self.__tf - self.__t0):
            dt0 = -dt0
        return dt0

    def _reset_state(self):
        """Back to (t0, y0, dt0) with zero counters.  Only allocates; the fills happen in `_materialize()` or, for
        a fused run, inside the kernel."""
        torch = _torch()
        n, dev = self.n_traj, self.device
        if self.__dict__.get("_hist_alias") and not self.host_io:
            # a history row handed out earlier (a[-1]) still refers to the state buffers: leave them to their
            # holders and continue on fresh ones (an allocation, not a copy)
            self._raw_y = torch.empty_like(self._y0)
            self._raw_t = torch.empty((n,), dtype=torch.float64, device=dev)
        if getattr(self, "_raw_y", None) is None:
            self._raw_y = torch.empty_like(self._y0)
            self._raw_t = torch.empty((n,), dtype=torch.float64, device=dev)
            self._raw_dt = torch.empty((n,), dtype=torch.float64, device=dev)
            self._raw_accepted = torch.empty(n, dtype=torch.int32, device=dev)
            self._raw_rejected = torch.empty(n, dtype=torch.int32, device=dev)
            self._raw_status = torch.empty(n, dtype=torch.int32, device=dev)
            self._counters = torch.zeros(8, dtype=torch.int64, device=dev)
        self._pristine = True
        self._stats = None
        self._pending = False
        self._event_log = None
        self._hist_alias = False
        self._hist_y = [self._y0]                         # row 0 = initial condition (never written to)
        self._hist_t = [self.__t0]                        # scalars stand for "every trajectory at this time"
        self._node_store = None
        self._steps_cache = None
        self._rows_used = 0
        self._protocol_path = False

    def _materialize(self):
        if self.__dict__.get("_pristine"):
            self._pristine = False
            self._raw_y.copy_(self._y0)
            self._raw_t.fill_(self.__t0)
            self._raw_dt.fill_(self._initial_dt())
            if self._raw_accepted is not None:
                self._raw_accepted.zero_()
            if self._raw_rejected is not None:
                self._raw_rejected.zero_()
            if self._raw_status is not None:
                self._raw_status.zero_()

    def _detach_history(self):
        if self.__dict__.get("_hist_alias"):
            self._hist_alias = False
            if not self.host_io and self._hist_y and self._hist_y[-1] is self._raw_y:
                # the history row (and whoever was handed a[-1]) keeps the old buffers; the state moves to copies
                self._raw_y = self._raw_y.clone()
                self._raw_t = self._raw_t.clone()

    def _hist_time(self, row):
        torch = _torch()
        if isinstance(row, float):
            row = torch.full((self.n_traj,), row, dtype=torch.float64, device="cpu" if self.host_io else self.device)
        return row

    def _fix_dt_dir(self, t1, t0):
        # in place: the stage-level kernels keep raw pointers to the dt buffer for the whole of integrate()
        torch = _torch()
        dt = self._dt
        flip = torch.sign(dt) != float(np.sign(t1 - t0))
        dt.copy_(torch.where(flip, -dt, dt))

    # ------------------------------------------------------------------ properties
    @property
    def sol(self):
        return DenseOutput(self) if self.__dense_output else None

    @property
    def success(self):
        self._resolve_pending()
        return self.__int_status == 1 or self.__int_status == 2

    def _pending_counters(self):
        """Device int64[4] = the fused kernel's [attempts, accepted, failed, unfinished] of the last lazy integrate()
        call (None if there is none): what a sharded run all-reduces without a host round trip."""
        return self._counters[2:6] if self._pending else None

    def _resolve_pending(self):
        """lazy_status: read the counters of the last launch now (one 64-byte device-to-host copy = the first host
        synchronisation since integrate() was called) and settle `stats` / `integration_status`."""
        if not self.__dict__.get("_pending"):
            return
        self._pending = False
        st = self._fused_stats()
        prev = self._stats or dict(attempts=0, accepted=0)
        self._stats = dict(attempts=prev["attempts"] + st[0], accepted=prev["accepted"] + st[1], failed=st[2], running=st[3])
        if st[2]:
            new_e = etypes.FailedIntegration("Failed to integrate system")
            new_e.__cause__ = etypes.FailedToMeetTolerances(
                "Failed to integrate {} of {} trajectories to the tolerances required: rtol={}, atol={}".format(
                    st[2], self.n_traj, self.integrator.rtol, self.integrator.atol))
            self.__int_status = new_e
        else:
            self.__int_status = 1

    def check(self):
        """Raises the `FailedIntegration` a lazy integrate() call could not raise itself (lazy_status=True)."""
        self._resolve_pending()
        if isinstance(self.__int_status, etypes.FailedIntegration):
            raise self.__int_status

    @property
    def events(self):
        """Events found so far.  A single system gives the reference's list of `StateTuple(t, y, event)`
        (`differential_system.py:557-560`); an ensemble gives an `events.EventLog` (per-trajectory counts, times,
        states and event indices as device tensors; `log.for_trajectory(i)` is the reference's list)."""
        if self._event_log is None:
            return self.__events
        return self._event_log if self.ensemble else self._event_log.for_trajectory(0)

    def _fire_callbacks(self, callbacks):
        """User callbacks see the system as it stands NOW: whatever was cached from the node store is dropped first
        (`a.t[-1]`, `a.y[-1]`, `a.sol` inside a callback; reference differential_system.py:1073-1084)."""
        self._steps_cache = None
        if self._node_store is not None:
            self._node_store.invalidate()
        for cbk in callbacks:
            cbk(self)

    def _step_history(self):
        """(t[steps + 1], y[steps + 1, *dim]) of a single system: every accepted step, row 0 = the initial condition --
        gathered from the node store by one kernel (dsb_node_gather) and cached until the next integrate() call."""
        torch = _torch()
        if self._steps_cache is None:
            st = self._node_store
            if st is None or int(st.node_count[0].item()) == 0:
                self._materialize()
                self._steps_cache = (self._raw_t.reshape(1).clone(), self._raw_y.reshape((1,) + tuple(self.dim)).clone())
            else:
                t_rows, y_rows = st.gather(0)
                self._steps_cache = (t_rows, y_rows.reshape((t_rows.shape[0],) + tuple(self.dim)))
        return self._steps_cache

    @staticmethod
    def _locate_step(t_rows, value):
        """The reference's `search_bisection` (utilities/utilities.py:223-267): index of the first row whose time is
        >= value, clamped to the stored range (rows in descending time for a backward integration are handled by
        mirroring)."""
        t_host = t_rows.cpu().numpy()
        rows = t_host.shape[0]
        asc = rows < 2 or t_host[-1] >= t_host[0]
        arr, val = (t_host, float(value)) if asc else (-t_host, -float(value))
        return int(min(np.searchsorted(arr, val, side="left"), rows - 1))

    def steps(self, trajectory=0):
        """(t, y) of every accepted step of ONE trajectory (row 0 = its initial condition): the reference's `a.t`,
        `a.y` for that trajectory.  Needs the node store (dense_output=True, or a single system)."""
        if self._node_store is None:
            raise ValueError("no step history: construct the system with dense_output=True (ensembles) and integrate first")
        t_rows, y_rows = self._node_store.gather(int(trajectory))
        return t_rows, y_rows.reshape((t_rows.shape[0],) + tuple(self.dim))

    @property
    def y(self):
        """History of states.  Single system: one row per accepted step like the reference (`:1015`), shape
        (steps + 1, *dim).  Ensemble (or history="calls"): one row per integrate() call (row 0 = y0), (rows, *dim[, N])."""
        torch = _torch()
        if self._record_steps:
            return self._step_history()[1]
        return torch.stack([self._user_shape(r) for r in self._hist_y])

    @property
    def t(self):
        torch = _torch()
        if self._record_steps:
            return self._step_history()[0]
        return torch.stack([self._user_time(self._hist_time(r)) for r in self._hist_t])

    @property
    def nfev(self):
        """Right-hand-side evaluations in the reference's accounting, summed over the ensemble:
        per trajectory 1 (constructor) + 1 (first initial_rhs) + (S+1) per attempt
        (S per attempt for FSAL schemes); SURVEY.md fact 5."""
        # the constructor's evaluation is counted until reset() zeroes the counter (reference :906): a reset system
        # starts again from 0, not from 1
        ctor = 0 if self.__dict__.get("_nfev_reset") else 1
        if self.__dict__.get("_pristine"):
            return self.n_traj * ctor
        if self.__dict__.get("_protocol_path"):
            return int(self.equ_rhs.nfev)                      # the protocol loop calls the wrapped callable itself
        att = int((self.accepted.sum() + self.rejected.sum()).item())
        integ = self.integrator
        per_attempt = integ.stages + (0 if getattr(integ, "is_fsal", False) or integ.symplectic else 1)
        frozen = self.__dict__.get("_nfev_frozen")
        if frozen is not None:                                 # evaluations of earlier integrator objects + this one's
            done, att0 = frozen
            return done if att == att0 else done + self.n_traj + per_attempt * (att - att0)
        if att == 0:
            return self.n_traj * ctor
        return self.n_traj * (ctor + 1) + per_attempt * att

    @property
    def njev(self):
        return 0

    @property
    def stats(self):
        """Ensemble totals since the last reset: step attempts, accepted steps, trajectories that failed the
        tolerances, trajectories still short of tf.  After a fused run they come from the kernel's own counters
        (no reduction launches); otherwise from the per-trajectory tensors."""
        if self.__dict__.get("_pristine"):
            return dict(attempts=0, accepted=0, failed=0, running=0)
        self._resolve_pending()
        if self._stats is not None:
            return dict(self._stats)
        acc = int(self.accepted.sum().item())
        return dict(attempts=acc + int(self.rejected.sum().item()), accepted=acc,
                    failed=int((self.status == cb.TRAJ_TOL_FAIL).sum().item()),
                    running=int((self.status == cb.TRAJ_RUNNING).sum().item()))

    @property
    def constants(self):
        return self.__consts

    @constants.setter
    def constants(self, new_constants):
        self.__consts = self._device_constants(new_constants)

    @constants.deleter
    def constants(self):
        self.__consts = dict()

    @property
    def rtol(self):
        return self.__rtol

    @rtol.setter
    def rtol(self, new_rtol):
        self.__rtol = new_rtol if np.size(new_rtol) > 1 else float(new_rtol)
        self.initialise_integrator()

    @property
    def atol(self):
        return self.__atol

    @atol.setter
    def atol(self, new_atol):
        self.__atol = new_atol if np.size(new_atol) > 1 else float(new_atol)
        self.initialise_integrator()

    @property
    def dt(self):
        """Current step size(s): a 0-d tensor for a single system, shape (N,) for an ensemble."""
        return self._user_time(self._dt)

    @dt.setter
    def dt(self, new_dt):
        # written IN PLACE: a callback that assigns `ode_sys.dt` (e.g. solve_ivp's max_step / min_step clamp,
        # reference :1229-1232) must reach kernels that hold a raw pointer to this buffer
        torch = _torch()
        new_dt = torch.as_tensor(new_dt, dtype=torch.float64, device=self._dt.device)
        if new_dt.ndim > 1 or (new_dt.ndim == 1 and new_dt.shape[0] not in (1, self.n_traj)):
            raise ValueError("dt must be a scalar or have shape (%d,)" % self.n_traj)
        self._dt.copy_(new_dt.reshape(-1).expand(self.n_traj) if new_dt.ndim else new_dt.expand(self.n_traj))
        self._fix_dt_dir(self.__tf, self.__t0)

    @property
    def t0(self):
        return self.__t0

    @t0.setter
    def t0(self, new_t0):
        if abs(self.__tf - float(new_t0)) <= D.epsilon("float64"):
            raise ValueError("The start time of the integration cannot be greater than or equal to {}!".format(self.__tf))
        self.__t0 = float(new_t0)
        self._fix_dt_dir(self.__tf, self.__t0)

    @property
    def tf(self):
        return self.__tf

    @tf.setter
    def tf(self, new_tf):
        if abs(self.__t0 - float(new_tf)) <= D.epsilon("float64"):
            raise ValueError("The end time of the integration cannot be equal to the start time: {}!".format(self.__t0))
        self.__tf = float(new_tf)
        self._fix_dt_dir(self.__tf, self.__t0)

    def get_current_time(self):
        return self._user_time(self._t)

    @property
    def counter(self):
        """Index of the last stored row: `a.t[a.counter]` is the current time (reference `:523, :792`)."""
        return len(self) - 1

    @property
    def events_dict(self):
        """{event function: StateTuple(t = its times, y = its states, event)} of a single system (reference `:573-593`)."""
        torch = _torch()
        if self.ensemble:
            # an ensemble's records live in ONE EventLog (per-trajectory counts, times, states, event indices on the
            # device): every event function maps to it; `log.for_trajectory(i)` is the reference's list for trajectory i
            log = self.events
            return {ev: log for ev in getattr(log, "events", [])}
        found = list(self.events)
        out = {}
        for ev in {id(e.event): e.event for e in found}.values():
            mine = [e for e in found if e.event is ev]
            out[ev] = StateTuple(t=torch.stack([torch.as_tensor(e.t) for e in mine]), y=torch.stack([e.y for e in mine]), event=ev)
        return out

    def get_step_interpolant(self):
        """(t_end, interpolant) of the last step (reference `:871-872` = `integrator.dense_output()`): from the integrator
        when the protocol loop drove it, else from the last two nodes of the store (dense_output=True)."""
        if self.__dict__.get("_protocol_path") or self._node_store is None or not self._node_store.with_f:
            return self.integrator.dense_output()
        from .utilities.interpolation import CubicHermiteInterp
        if self.ensemble:
            raise NotImplementedError("per-step interpolants of an ensemble are served by `a.sol` (one launch for all trajectories)")
        count = int(self._node_store.node_count[0].item())
        if count < 2:
            return self.integrator.dense_output()
        t_rows, y_rows, f_rows = self._node_store.gather(0, k0=count - 2, count=2, with_f=True)
        shape = tuple(self.dim)
        interp = CubicHermiteInterp(t_rows[0], t_rows[1], y_rows[0].reshape(shape), y_rows[1].reshape(shape),
                                    f_rows[0].reshape(shape), f_rows[1].reshape(shape))
        return t_rows[1], interp

    @property
    def staggered_mask(self):
        """The kick mask of the splitting schemes (reference `:531, :820-823`): the default -- second half of axis 0 --
        is the only one (a custom mask is dead code in the reference)."""
        return getattr(self.integrator, "staggered_mask", None)

    def set_kick_vars(self, staggered_mask):
        """reference `:769-788`.  Only the default mask is supported."""
        if staggered_mask is not None:
            raise NotImplementedError("custom staggered_mask is not supported (dead code in the reference)")
        self.initialise_integrator()

    # ------------------------------------------------------------------ method handling
    def initialise_integrator(self, preserve_states=False):
        kwargs = dict(dtype=self.dtype, device=self.device, atol=self.atol, rtol=self.rtol)
        if self.__dict__.get("integrator") is not None and not self.__dict__.get("_pristine", True) \
                and not self.__dict__.get("_protocol_path"):
            # a NEW integrator object in the middle of a run (method switch, new tolerances; reference :646, :660, :869):
            # what the old one evaluated is final, and the new one evaluates initial_rhs again on its first step
            # (integrator_types.py:169-175: final_rhs is None)
            self.__dict__["_nfev_frozen"] = (self.nfev, int((self.accepted.sum() + self.rejected.sum()).item()))
        self.integrator = self.__method(self.dim, **kwargs)

    @property
    def method(self):
        return self.__method

    @method.setter
    def method(self, new_method):
        self.set_method(new_method, None)

    def set_method(self, new_method, staggered_mask=None, preserve_states=True):
        if staggered_mask is not None:
            raise NotImplementedError("custom staggered_mask is not supported (dead code in the reference)")
        if isinstance(new_method, str) and new_method in integrators.available_methods():
            self.__method = integrators.available_methods(False)[new_method]
        elif isinstance(new_method, type) and issubclass(new_method, integrators.IntegratorTemplate):
            self.__method = new_method
        else:
            if isinstance(new_method, str) and (new_method in integrators.REFERENCE_ONLY_METHODS or "Richardson" in new_method):
                raise NotImplementedError(
                    "{!r} is an implicit scheme of the reference; desolver_b200 replaces the explicit Runge-Kutta and "
                    "symplectic hot path only (implicit_methods() is empty) -- use desolver itself for it".format(new_method))
            raise ValueError("The method you selected does not exist in the list of available methods, "
                             "call desolver_b200.available_methods() to see what these are")
        self.initialise_integrator(preserve_states=preserve_states)

    @property
    def integration_status(self):
        self._resolve_pending()
        if self.__int_status == 0:
            return "Integration has not been run."
        if isinstance(self.__int_status, KeyboardInterrupt):
            return "A KeyboardInterrupt exception was raised during integration."
        if isinstance(self.__int_status, etypes.FailedIntegration):
            if self.__int_status.__cause__ is not None:
                return "The integration failed with the following exception: {}\n\tCaused by {}".format(
                    self.__int_status, self.__int_status.__cause__)
            return "The integration failed with the following exception: {}".format(self.__int_status)
        if self.__int_status == 1:
            return "Integration completed successfully."
        if self.__int_status == 2:
            return "Integration terminated upon finding a triggered event."
        return "It should not be possible to initialise the system with this status"

    def reset(self):
        """Resets the system to the initial time (reference :900-910)."""
        self.__int_status = 0
        self.__events = []
        self.equ_rhs.nfev = 0                              # reference :906
        self.__dict__["_nfev_reset"] = True
        self.__dict__["_nfev_frozen"] = None
        self._reset_state()
        self.initialise_integrator()

    # ------------------------------------------------------------------ the hot path
    def _fused_plugin(self):
        """(plugin id, parameter names) if the RHS is tagged with a fused device plugin that the
        library has a thread-per-trajectory kernel for; else None."""
        rhs = self.equ_rhs
        pid = getattr(rhs, "plugin_id", None)
        if pid is None or getattr(rhs, "device_plugin", None) not in rhs_plugins.FUSED:
            return None
        if not isinstance(self.integrator, integrators.RungeKuttaIntegrator):
            return None
        if self.integrator.stages not in cb.FUSED_STAGES or len(self.dim) != 1:
            return None
        if self.integrator.has_vector_tolerances:          # one tolerance per component: stage-level kernels
            return None
        if self.strict and str(rhs.device_plugin).startswith("user:"):
            return None                                     # caller-built plug-ins carry the fast flavour only
        return pid, tuple(rhs.plugin_params), dict(rhs.plugin_defaults)

    def _param_tensors(self, names, defaults):
        """Per-plugin parameter arrays: a (N,) tensor is per-trajectory (stride 1), anything else
        is one shared value (stride 0).  Scalars are uploaded once and cached (a fresh `torch.as_tensor`
        per integrate() call costs a host-to-device copy each)."""
        torch = _torch()
        tensors, strides = [], []
        for name in names:
            v = self.__consts.get(name, defaults.get(name))
            if not isinstance(v, torch.Tensor):
                key = (name, float(v))
                cached = self._scalar_cache.get(key)
                if cached is None:
                    cached = self._scalar_cache[key] = torch.full((1,), float(v), dtype=torch.float64, device=self.device)
                tensors.append(cached)
                strides.append(0)
                continue
            v = v.to(device=self.device, dtype=torch.float64)
            if v.ndim == 0:
                tensors.append(v.reshape(1).contiguous())
                strides.append(0)
            elif tuple(v.shape) == (self.n_traj,):
                tensors.append(v.contiguous())
                strides.append(1)
            else:
                raise ValueError("constant %r must be a scalar or have shape (%d,)" % (name, self.n_traj))
        return tensors, strides

    @property
    def _want_nodes(self):
        return self.__dense_output or self._record_steps

    def _ensure_node_store(self, kind, need=0, chunked=False):
        """The node store of this system (dense output and/or single-system step history), created on first use.
        kind "fused": written by the fused kernel (32-entry groups, pages claimed per warp); kind "rows": written by
        dsb_node_append_rows between stage kernels (one row-group of all n trajectories per attempt).  `need` = entries
        (fused) / row-groups (rows) that must still be free."""
        torch = _torch()
        from .node_store import NodeStore
        if not self._want_nodes:
            return None
        n, st = self.n_traj, self._node_store
        if st is not None and st.kind != kind:
            if kind != "rows":
                # (not reached through integrate(): a system whose store is row-written stays on the stage-level path)
                raise NotImplementedError("this system's node store was started by the stage-level kernels; reset() before "
                                          "running it through the fused kernel")
            # the run began in the fused kernel and continues on the stage-level kernels (another method, per-component
            # tolerances, events): the same nodes, re-laid out for the row writer
            st = self._node_store = st.to_rows(with_f=bool(self.__dense_output))
            self._rows_used = int(st.cursor[0].item()) // n
        if st is None:
            if kind == "fused":
                page = 32 if (chunked or n <= 4096) else 256
                per_traj = self._dense_capacity if self._dense_capacity is not None else (5001 if n == 1 else 288)
                want = n * per_traj + min(4736, (n + 31) // 32) * page
                free = torch.cuda.mem_get_info(self.device)[0]
                cap = min(want, int(0.6 * free) // (8 * (2 + 2 * self.numel)))
                cap = max(cap, int(need), 2 * page)
                st = NodeStore(self.device, self.numel, n, "fused", cap, page_entries=page, with_f=True)
            else:
                rows = self._dense_capacity if self._dense_capacity is not None else 64
                st = NodeStore(self.device, self.numel, n, "rows", max(rows, int(need), 2) * n, with_f=self.__dense_output)
            self._node_store = st
        if kind == "rows" and need:
            if (self._rows_used + need) * n > st.capacity:
                st.grow(max(2 * st.capacity, (self._rows_used + need) * n))
        elif kind == "fused" and need:
            claimed = st.usage()[0]
            if claimed + need > st.capacity:
                st.grow(claimed + need)
        return st

    def _event_tensors(self, events):
        """Device-side description of the events (dsb_erk_run_args.ev_*): coefficient rows, flag words, record store."""
        torch = _torch()
        from .events import EventLog
        coef = torch.as_tensor(np.stack([ev.coefficients(self.numel) for ev in events]), dtype=torch.float64).to(self.device)
        flags = torch.as_tensor([ev.flags for ev in events], dtype=torch.int32).to(self.device)
        self._event_coef_dy = None
        if any(ev.requires_dstate and hasattr(ev, "w") for ev in events):          # weights on dState/dt
            self._event_coef_dy = torch.as_tensor(np.stack([ev.dy_coefficients(self.numel) for ev in events]),
                                                  dtype=torch.float64).to(self.device)
        # the log accumulates over integrate() calls like the reference's `self.__events` (:1043-1048); it is only
        # re-created when the SET of events changes -- compared by value (coefficients, flags), because a plain
        # callable is converted to a new HyperplaneEvent object on every call
        signature = tuple(ev.signature if hasattr(ev, "signature") else
                          (tuple(ev.coefficients(self.numel).tolist()), tuple(ev.dy_coefficients(self.numel).tolist()), int(ev.flags))
                          for ev in events)
        if self._event_log is None or self._event_log.compiled != signature:
            records = torch.zeros((self.n_traj, int(self.event_capacity), self.numel + 2), dtype=torch.float64, device=self.device)
            count = torch.zeros(self.n_traj, dtype=torch.int32, device=self.device)
            self._event_log = EventLog(events, records, count, self.dim)
            self._event_log.compiled = signature
        return coef, flags

    def _run_fused(self, tf, plugin, max_steps, first_call, events=None):
        """One dsb_erk_run launch.  With a node store (dense output / step history) the pool is sized before the launch:
        exactly for a chunked call (a trajectory records at most max_steps + 1 nodes), by estimate for a one-shot call --
        and if the estimate was too small, the kernel has still COUNTED every node, so the state is put back, the pool
        grown to the size that is now known, and the (bitwise reproducible) launch repeated."""
        torch = _torch()
        store = None
        if self._want_nodes and not events:
            n = self.n_traj
            chunked = int(max_steps) < self._step_cap
            store = self._ensure_node_store("fused", chunked=chunked)
            slack = min(4736, (n + 31) // 32) * store.page_entries         # one page in progress per warp
            if chunked:
                store = self._ensure_node_store("fused", need=n * (int(max_steps) + 1) + slack)
        if store is None or int(max_steps) < self._step_cap:
            self._launch_fused(tf, plugin, max_steps, first_call, events, store)
            if store is not None:
                self._steps_cache = None
            return
        pristine = bool(self._pristine and first_call)
        snap_store = store.snapshot()
        snap_state = None if pristine else tuple(x.clone() for x in (self._raw_y, self._raw_t, self._raw_dt, self._raw_accepted,
                                                                     self._raw_rejected, self._raw_status))
        nodes_before = int(snap_store[1].sum().item())
        for attempt in range(3):
            self._launch_fused(tf, plugin, max_steps, first_call, events, store)
            claimed, dropped = store.usage()
            if not dropped:
                break
            wanted = int(store.node_count.sum().item()) - nodes_before           # nodes this call records
            store.restore(snap_store)
            if pristine:
                self._pristine = True
            else:
                for dst, src in zip((self._raw_y, self._raw_t, self._raw_dt, self._raw_accepted, self._raw_rejected,
                                     self._raw_status), snap_state):
                    dst.copy_(src)
            store.grow(int(snap_store[0][0].item()) + wanted + wanted // 64 + slack)
        else:
            raise MemoryError("dense-output node store could not be sized")
        self._steps_cache = None

    def _launch_fused(self, tf, plugin, max_steps, first_call, events, store, stops=None):
        integ = self.integrator
        pid, names, defaults = plugin
        handle = cb.erk_handle(self.device.index or 0, pid, self.numel, integ.tableau_intermediate,
                               integ.tableau_final, integ.order, strict=self.strict)
        params, strides = self._param_tensors(names, defaults)
        self._resolve_pending()                   # a continued integration accumulates the previous call's totals
        self._counters.zero_()
        a = cb.ErkRunArgs()
        a.n = self.n_traj
        if self._pristine and first_call:
            # pristine state (constructor / reset()): the kernel reads y0 directly, starts every trajectory at
            # (t0, dt0) with zero counters and overwrites t, dt, accepted, rejected, status -- no fill launches
            self._pristine = False
            a.fresh, a.t_init, a.dt_init = 1, self.__t0, self._initial_dt()
            a.y_init = self._y0.data_ptr()
            sc = self._scatter_store
            if sc is not None and not events and store is None and int(max_steps) >= self._step_cap:
                # sharded ensemble: results go straight to slot i * W + rank of the root's gather buffer (peer memory);
                # the local y / t / dt / counter arrays are NOT written by this launch
                a.out_y, a.out_t, a.out_dt = sc["y"], sc["t"], sc["dt"]
                a.out_accepted, a.out_rejected, a.out_status = sc["accepted"], sc["rejected"], sc["status"]
                a.out_stride, a.out_index_mul, a.out_index_add = sc["stride"], sc["index_mul"], sc["index_add"]
        elif self._scatter_store is not None:
            raise RuntimeError("scatter_store systems integrate once per reset() (the local state is not kept)")
        a.y, a.t, a.dt = self._raw_y.data_ptr(), self._raw_t.data_ptr(), self._raw_dt.data_ptr()
        for i, (p, s) in enumerate(zip(params, strides)):
            a.params[i] = p.data_ptr()
            a.param_stride[i] = s
        a.tf, a.rtol, a.atol = float(tf), float(integ.rtol), float(integ.atol)
        a.accepted, a.rejected, a.status = (self._raw_accepted.data_ptr(), self._raw_rejected.data_ptr(),
                                            self._raw_status.data_ptr())
        a.max_steps = int(max_steps)
        a.first_call = 1 if first_call else 0
        if store is not None:
            a.store = store.struct()
            store.invalidate()
        if stops is not None:
            a.stops, a.n_stops = stops[0].data_ptr(), int(stops[0].numel())
            a.stop_y, a.stop_stride = stops[1].data_ptr(), self.n_traj
        a.counters = self._counters.data_ptr()
        if events:
            coef, flags = self._event_tensors(events)
            a.n_events, a.ev_capacity = len(events), self._event_log.capacity
            a.ev_coef, a.ev_flags = coef.data_ptr(), flags.data_ptr()
            a.ev_records, a.ev_count = self._event_log._records.data_ptr(), self._event_log.count.data_ptr()
            if self._event_coef_dy is not None:
                a.ev_coef_dy = self._event_coef_dy.data_ptr()
            params = list(params) + [coef, flags]             # keep alive until the launch has run
        ev = self.__dict__.get("_kernel_events")          # optional (start, end) CUDA events bracketing the launch
        if ev:
            ev[0].record()
        handle.run(a, cb.current_stream_ptr(self.device))
        if ev:
            ev[1].record()
        self._launches += 1
        self._keepalive = params

    def _run_host_pipeline(self, tf, plugin):
        """host_io=True: the ensemble is cut into `host_chunks` contiguous pieces; piece c is uploaded on a copy
        stream, integrated by its own dsb_erk_run launch (pointers offset into the [dim][N] staging buffers,
        `stride` = N, `fresh` = 1 so there is no fill launch) on one of two compute streams, and downloaded on a third
        stream.  Upload of piece c+1, kernel of piece c and download of piece c-1 overlap; two compute streams
        let the next kernel's blocks move in while the previous kernel's last trajectories drain."""
        torch = _torch()
        integ = self.integrator
        pid, names, defaults = plugin
        n, dev = self.n_traj, self.device
        handle = cb.erk_handle(dev.index or 0, pid, self.numel, integ.tableau_intermediate, integ.tableau_final,
                               integ.order, strict=self.strict)
        params, strides = self._param_tensors(names, defaults)
        chunks = max(1, min(self.host_chunks, n // 1024 if n >= 1024 else 1))
        bounds = [((n * c // chunks) // 32) * 32 for c in range(chunks)] + [n]
        if self._pipe_streams is None:
            self._pipe_streams = tuple(torch.cuda.Stream(device=dev) for _ in range(4))
        up, down, comp0, comp1 = self._pipe_streams
        cur = torch.cuda.current_stream(dev)
        self._counters.zero_()
        if self.host_pipeline == "streamed":
            # ONE persistent launch; the library overlaps the chunked uploads / downloads with it through stream
            # memory operations (dsb_erk_run_host)
            shift = max(10, int(np.ceil(np.log2(max(1, -(-n // max(1, self.host_chunks)))))))
            if self._pipe_flags.numel() < 2 + (-(-n // (1 << shift))):
                self._pipe_flags = torch.zeros(2 + (-(-n // (1 << shift))), dtype=torch.int32, device=dev)
            ha = cb.ErkHostArgs()
            a = ha.run
            a.n, a.stride = n, n
            a.y, a.t, a.dt = self._dev_y.data_ptr(), self._raw_t.data_ptr(), self._raw_dt.data_ptr()
            for i, (p, sp) in enumerate(zip(params, strides)):
                a.params[i], a.param_stride[i] = p.data_ptr(), sp
            a.tf, a.rtol, a.atol = float(tf), float(integ.rtol), float(integ.atol)
            a.accepted, a.rejected, a.status = (self._dev_accepted.data_ptr(), self._dev_rejected.data_ptr(),
                                                self._dev_status.data_ptr())
            a.max_steps, a.first_call = int(self._step_cap), 1
            a.fresh, a.t_init, a.dt_init = 1, self.t0, self._initial_dt()
            a.counters = self._counters.data_ptr()
            ha.host_y0, ha.host_y = self._y0.data_ptr(), self._raw_y.data_ptr()
            ha.host_accepted = self._host_acc_buffer.data_ptr() if self._host_acc_buffer is not None else None
            ha.host_rejected = self._host_rej_buffer.data_ptr() if self._host_rej_buffer is not None else None
            ha.host_status = self._host_status_buffer.data_ptr() if self._host_status_buffer is not None else None
            ha.host_stride, ha.chunk_shift = n, shift
            ha.flags = self._pipe_flags.data_ptr()
            ha.stream_up, ha.stream_down = up.cuda_stream, down.cuda_stream
            if handle.run_host(ha, C.c_void_p(cur.cuda_stream)):
                self._launches += 1
                self._keepalive = params
                self._pristine = False
                c = self._counters.cpu()                      # synchronises `cur`, which waited for the download stream
                if int(c[6]):
                    raise RuntimeError("dsb_erk_run_host: the kernel timed out waiting for its initial states")
                self._raw_status = self._host_status_buffer
                self._raw_accepted, self._raw_rejected = self._host_acc_buffer, self._host_rej_buffer
                self._host_status_clean = int(c[4]) + int(c[5]) == 0
                return int(c[2]), int(c[3]), int(c[4]), int(c[5])
            self.host_pipeline = "chunked"                    # driver without stream memory operations
        for st in (up, comp0, comp1):
            st.wait_stream(cur)
        t0, dt0 = self.t0, self._initial_dt()
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            if hi <= lo:
                continue
            with torch.cuda.stream(up):
                for d in range(self.numel):
                    self._dev_y[d, lo:hi].copy_(self._y0[d, lo:hi], non_blocking=True)
                uploaded = up.record_event()
            comp = comp0 if c % 2 == 0 else comp1
            comp.wait_event(uploaded)
            a = cb.ErkRunArgs()
            a.n, a.stride = hi - lo, n
            a.y = self._dev_y.data_ptr() + 8 * lo
            a.t, a.dt = self._raw_t.data_ptr() + 8 * lo, self._raw_dt.data_ptr() + 8 * lo
            for i, (p, sp) in enumerate(zip(params, strides)):
                a.params[i] = p.data_ptr() + 8 * lo * sp
                a.param_stride[i] = sp
            a.tf, a.rtol, a.atol = float(tf), float(integ.rtol), float(integ.atol)
            a.accepted = self._dev_accepted.data_ptr() + 4 * lo
            a.rejected = self._dev_rejected.data_ptr() + 4 * lo
            a.status = self._dev_status.data_ptr() + 4 * lo
            a.max_steps, a.first_call = int(self._step_cap), 1
            a.fresh, a.t_init, a.dt_init = 1, t0, dt0
            a.counters = self._counters.data_ptr() + 64 * c
            handle.run(a, C.c_void_p(comp.cuda_stream))
            self._launches += 1
            integrated = comp.record_event()
            down.wait_event(integrated)
            with torch.cuda.stream(down):
                for d in range(self.numel):
                    self._raw_y[d, lo:hi].copy_(self._dev_y[d, lo:hi], non_blocking=True)
                if self._host_acc_buffer is not None:
                    self._host_acc_buffer[lo:hi].copy_(self._dev_accepted[lo:hi], non_blocking=True)
                if self._host_rej_buffer is not None:
                    self._host_rej_buffer[lo:hi].copy_(self._dev_rejected[lo:hi], non_blocking=True)
                if self._host_status_buffer is not None:
                    self._host_status_buffer[lo:hi].copy_(self._dev_status[lo:hi], non_blocking=True)
        for st in (down, comp0, comp1):
            cur.wait_stream(st)
        self._keepalive = params
        self._pristine = False
        c = self._counters.cpu().reshape(-1, 8)[:chunks]          # synchronises: every copy above has landed
        self._raw_status = self._host_status_buffer
        self._raw_accepted, self._raw_rejected = self._host_acc_buffer, self._host_rej_buffer
        self._host_status_clean = int(c[:, 4].sum()) + int(c[:, 5].sum()) == 0
        return int(c[:, 2].sum()), int(c[:, 3].sum()), int(c[:, 4].sum()), int(c[:, 5].sum())

    def integrate_stops(self, times):
        """Integrates every trajectory through the stop times `times` (monotone in the direction of integration) exactly
        as successive `integrate(t)` calls would -- the reference's `t_eval` loop (`differential_system.py:1245-1248`) --
        and returns the states at the stops, shape (T, *dim[, N]).  For a fused right-hand side with an explicit method
        of at most 7 stages this is ONE launch of the multi-stop kernel (dsb_erk_run_args.stops), bit-identical to the
        sequence of calls; otherwise it is that sequence."""
        torch = _torch()
        times = [float(x) for x in np.asarray(times, dtype=np.float64).reshape(-1)]
        if not times:
            raise ValueError("integrate_stops needs at least one time")
        plugin = self._fused_plugin() if self.use_fused else None
        fast = (plugin is not None and not self.host_io and not self._want_nodes and not self.integrator.symplectic
                and self.integrator.stages <= 7 and not getattr(self.integrator, "has_custom_adaptation", False))
        if not fast:
            rows = []
            for tq in times:
                self.integrate(tq)
                rows.append(self[-1].y.clone())
            return torch.stack(rows)
        with torch.cuda.device(self.device):
            self._detach_history()
            stops = torch.as_tensor(times, dtype=torch.float64).to(self.device)
            out = torch.empty((len(times), self.numel, self.n_traj), dtype=torch.float64, device=self.device)
            self._launch_fused(times[-1], plugin, -1, True, None, None, stops=(stops, out))
            st = self._fused_stats()
            prev = self._stats or dict(attempts=0, accepted=0)
            self._stats = dict(attempts=prev["attempts"] + st[0], accepted=prev["accepted"] + st[1], failed=st[2], running=st[3])
            for k in range(len(times) - 1):                        # one history row per stop, like one per integrate() call
                self._hist_y.append(out[k])
                self._hist_t.append(times[k])
            self._hist_y.append(self._raw_y)
            self._hist_t.append(self._raw_t)
            self._hist_alias = True
            self._keepalive = (self._keepalive, stops, out)
            if st[2]:
                new_e = etypes.FailedIntegration("Failed to integrate system")
                new_e.__cause__ = etypes.FailedToMeetTolerances(
                    "Failed to integrate {} of {} trajectories to the tolerances required: rtol={}, atol={}".format(
                        st[2], self.n_traj, self.integrator.rtol, self.integrator.atol))
                self.__int_status = new_e
                raise new_e
            self.__int_status = 1
        return out.reshape((len(times),) + tuple(self.dim) + ((self.n_traj,) if self.ensemble else ()))

    def _run_push(self, tf, plugin):
        """Sharded ensemble, gather overlapped with the integration: ONE flow-controlled persistent launch
        (dsb_erk_run_host without uploads) whose download stream copies chunk c of this shard's results to the ROOT rank's
        staging buffer (peer memory, CUDA IPC) as soon as the chunk's trajectories are done.  The local state arrays are
        written as usual.  Nothing is read back: the caller all-reduces the kernel's counters (lazy status)."""
        torch = _torch()
        integ = self.integrator
        pid, names, defaults = plugin
        n, dev = self.n_traj, self.device
        handle = cb.erk_handle(dev.index or 0, pid, self.numel, integ.tableau_intermediate, integ.tableau_final, integ.order,
                               strict=self.strict)
        params, strides = self._param_tensors(names, defaults)
        tgt = self._push_target
        if self.__dict__.get("_push_streams") is None:
            self._push_streams = torch.cuda.Stream(device=dev)
        down = self._push_streams
        cur = torch.cuda.current_stream(dev)
        shift = max(10, int(np.ceil(np.log2(max(1, -(-n // max(1, self.push_chunks)))))))
        nflags = 2 + (-(-n // (1 << shift)))
        if self.__dict__.get("_push_flags") is None or self._push_flags.numel() < nflags:
            self._push_flags = torch.zeros(nflags, dtype=torch.int32, device=dev)
        self._counters.zero_()
        ha = cb.ErkHostArgs()
        a = ha.run
        a.n, a.stride = n, n
        a.y, a.t, a.dt = self._raw_y.data_ptr(), self._raw_t.data_ptr(), self._raw_dt.data_ptr()
        a.y_init = self._y0.data_ptr()
        for i, (p, sp) in enumerate(zip(params, strides)):
            a.params[i], a.param_stride[i] = p.data_ptr(), sp
        a.tf, a.rtol, a.atol = float(tf), float(integ.rtol), float(integ.atol)
        a.accepted, a.rejected, a.status = self._raw_accepted.data_ptr(), self._raw_rejected.data_ptr(), self._raw_status.data_ptr()
        a.max_steps, a.first_call = int(self._step_cap), 1
        a.fresh, a.t_init, a.dt_init = 1, self.t0, self._initial_dt()
        a.counters = self._counters.data_ptr()
        ha.host_y0 = None
        ha.host_y, ha.host_t = tgt["y"], tgt["t"]
        ha.host_accepted, ha.host_rejected, ha.host_status = tgt["accepted"], tgt["rejected"], tgt["status"]
        ha.host_stride, ha.chunk_shift = int(tgt["n_max"]), shift
        ha.flags = self._push_flags.data_ptr()
        ha.stream_up, ha.stream_down = None, down.cuda_stream
        if not handle.run_host(ha, C.c_void_p(cur.cuda_stream)):
            raise RuntimeError("the pushed gather needs the flow-controlled kernel and stream memory operations "
                               "(dsb_erk_run_host returned DSB_ERR_UNSUPPORTED)")
        self._launches += 1
        self._keepalive = params
        self._pristine = False

    def _fused_stats(self):
        """(attempts, accepted, failed, running) of the last dsb_erk_run, from the kernel's own counters: ONE
        64-byte device-to-host copy instead of reductions over the per-trajectory arrays."""
        c = self._counters.cpu()
        return int(c[2]), int(c[3]), int(c[4]), int(c[5])

    def _run_stagewise(self, tf, callbacks, bar, events=None):
        """Generic right-hand side: the user's torch callable is evaluated between the fused stage
        kernels of libdesolver_b200.so (dsb_erk_stage_*), all on the current stream.  The host looks at
        the active-trajectory count only every `self.sync_every` attempts.  When fewer than half of the
        working columns are still active (trajectories need very different numbers of attempts, e.g. Van der
        Pol with mu in [0.1, 10]) the working set is compacted to the active trajectories, so the lock-step
        rounds near the end no longer pay for the finished ones."""
        torch = _torch()
        integ = self.integrator
        if not isinstance(integ, integrators.RungeKuttaIntegrator):
            raise NotImplementedError("stage-level path supports explicit Runge-Kutta methods")
        n, D, S, dev = self.n_traj, self.numel, integ.stages, self.device
        # a right-hand side may keep derived device data (rhs_plugins.linear: the INT8 digit planes of A); it refreshes them
        # here, eagerly, because the attempts below are replayed as CUDA graphs that only see buffers, not Python
        prepare = getattr(self.equ_rhs, "prepare", None)
        if prepare is not None:
            prepare(**self.__consts)
        handle = cb.erk_handle(dev.index or 0, 100, D, integ.tableau_intermediate, integ.tableau_final, integ.order,
                               strict=self.strict)
        chunks = max(1, min(4096, (D + 127) // 128))
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        stream = cb.current_stream_ptr(dev)
        store = self._ensure_node_store("rows", need=2) if self._want_nodes else None
        dense = store is not None
        full = dict(y=self._y, t=self._t, dt=self._dt, accepted=self.accepted, rejected=self.rejected, status=self.status)
        persistent = ("h", "h0", "trial", "flags", "scale", "hist")     # per-trajectory state that outlives an attempt
        tol_vec = integ.tolerance_arrays(dev) if integ.has_vector_tolerances else None

        def alloc(m):
            return dict(h=torch.zeros(m, **f64), h0=torch.zeros(m, **f64), y_stage=torch.empty((D, m), **f64),
                        t_stage=torch.zeros(m, **f64), dY=torch.zeros((D, m), **f64), scale=torch.zeros((D, m), **f64),
                        hist=torch.ones((3, m), **f64), partial=torch.zeros((chunks, m), **f64),
                        trial=torch.zeros(m, **i32), flags=torch.zeros(m, **i32))

        def bind(work, buf, m):
            a = cb.ErkStageArgs()
            a.n, a.dim, a.chunks = m, D, chunks
            a.y, a.t, a.dt = work["y"].data_ptr(), work["t"].data_ptr(), work["dt"].data_ptr()
            for k, v in buf.items():
                setattr(a, k, v.data_ptr())
            a.tf = float(tf)
            if tol_vec is not None:
                a.rtol_vec, a.atol_vec = tol_vec[0].data_ptr(), tol_vec[1].data_ptr()
            else:
                a.rtol, a.atol = float(integ.rtol), float(integ.atol)
            a.accepted, a.rejected, a.status = work["accepted"].data_ptr(), work["rejected"].data_ptr(), work["status"].data_ptr()
            a.counters = self._counters.data_ptr()
            shape = tuple(self.dim) + ((m,) if self.ensemble else ())
            return a, buf["y_stage"].reshape(shape), (buf["t_stage"] if self.ensemble else buf["t_stage"].reshape(()))

        # persistent stage arrays for right-hand sides with an `out=` argument (the linear plugin: 13 x 268 MB for the
        # benchmark configuration, which would otherwise be allocated by every eager attempt AND inside every graph's
        # private pool, and released again at the end of every integrate() call)
        k_out = None
        if getattr(self.equ_rhs, "supports_out", False) and self.ensemble:
            cache = self.__dict__.get("_stage_k_cache")
            if cache is None or len(cache) != S or tuple(cache[0].shape) != (D, n):
                cache = self._stage_k_cache = [torch.empty((D, n), **f64) for _ in range(S)]
            k_out = cache
        work, buf, m, index, consts = full, alloc(n), n, None, self.__consts
        m_live = n
        a, y_view, t_view = bind(work, buf, m)
        search = None
        if events:
            # device event detection around any right-hand side (dsb_erk_stage_finish_events): per-trajectory final
            # times (a terminal event re-targets its trajectory), scratch state, and the record store of the log
            coef, flags = self._event_tensors(events)
            ev_state = torch.zeros((2 + 2 * cb.MAX_STAGE_EVENTS, n), **f64)
            tf_per = torch.full((n,), float(tf), **f64)
            a.n_events, a.ev_capacity = len(events), self._event_log.capacity
            a.ev_coef, a.ev_flags = coef.data_ptr(), flags.data_ptr()
            a.ev_records, a.ev_count = self._event_log._records.data_ptr(), self._event_log.count.data_ptr()
            a.ev_state, a.tf_per = ev_state.data_ptr(), tf_per.data_ptr()
            if self._event_coef_dy is not None:
                a.ev_coef_dy = self._event_coef_dy.data_ptr()
            self._keepalive = (coef, flags, ev_state, tf_per)      # referenced by raw pointer until the run has finished
            if any(not hasattr(ev, "w") for ev in events):
                # arbitrary event callables: their values come from the caller's code between the event-search kernels
                from .events import EventSearch
                search = self._event_search = EventSearch(self, events, self.__consts)
                f_start = None
                if search.needs_dy:
                    f_start = self.equ_rhs(self._user_time(self._t), self._user_shape(self._y), **self.__consts).to(torch.float64).reshape(D, n).contiguous()
                search.start(self._t, self._y, f_start)
        can_compact = self.ensemble and self.compaction and not callbacks and not dense and not events
        K = [None] * S
        y_committed = self._user_shape(self._y)
        t_committed = self._user_time(self._t)

        def record_nodes(flags_ptr):
            """node (t, y[, f = rhs(t, y)]) for the trajectories flagged as just accepted (all if flags_ptr is None); f only
            with dense output -- the step history of a single system needs (t, y) alone and costs no evaluation"""
            f = None
            if store.with_f:
                f = self.equ_rhs(t_committed, y_committed, **self.__consts).to(torch.float64).reshape(D, n)
                f = f if f.is_contiguous() else f.contiguous()
                keep_node_f[0] = f
            store.append_rows(self._t, self._y, f, flags_ptr, 4)
            self._rows_used += 1
            self._launches += 2

        keep_node_f = [None]

        def scatter_back():
            for k, v in work.items():                   # (padding columns beyond m_live are copies: not written back)
                full[k].index_copy_(full[k].ndim - 1, index[:m_live], v[..., :m_live])

        if dense and int(store.node_count.max().item()) == 0:
            record_nodes(None)                                   # node 0 = initial condition
        first, attempts = True, 0
        acc_seen = int(work["accepted"].sum().item()) if (callbacks and not self.ensemble) else 0
        # heavy attempts (large state x ensemble): one host look per attempt costs nothing next to the
        # attempt itself and saves up to sync_every-1 wasted lock-step rounds at the end
        sync_every = 1 if D * n * S >= (1 << 24) else self.sync_every

        def one_attempt(first_call):
            """begin, S x (stage state -> user RHS), finish[, dense nodes]: ~S*(1 + RHS kernels) + 4 launches"""
            st = cb.current_stream_ptr(dev)
            handle.stage_begin(a, first_call, st)
            for s in range(S):
                handle.stage_state(a, s, st)
                if k_out is not None:
                    # right-hand side that writes into a caller-owned array (rhs.supports_out): the S stage arrays are
                    # owned by the system and reused by every attempt, graph and integrate() call -- no allocation
                    k = self.equ_rhs(t_view, y_view, out=k_out[s].reshape(-1)[:D * m].reshape(D, m), **consts)
                else:
                    k = self.equ_rhs(t_view, y_view, **consts)
                k = k.to(torch.float64).reshape(D, m)
                K[s] = k if k.is_contiguous() else k.contiguous()      # keep alive until the attempt finishes
                a.K[s] = K[s].data_ptr()
            if events:
                # phase 0: controller + (t_stage, y_stage) = the state after the attempt; f1 = rhs there (the reference's
                # final_rhs, integrator_types.py:308); phase 1: locate / record / roll back, then the masked commit
                handle.stage_finish_events(a, 0, None, st)
                f1 = self.equ_rhs(t_view, y_view, **consts).to(torch.float64).reshape(D, m)
                f1 = f1 if f1.is_contiguous() else f1.contiguous()
                K.append(f1)                                           # keep alive until the attempt finishes
                del K[S:-1]
                if search is not None:
                    roots = search.locate(buf["flags"], 4, 8, ev_state[0], work["t"], work["y"], buf["y_stage"], K[0], f1)
                    handle.stage_finish_events_located(a, f1.data_ptr(), roots.data_ptr(), st)
                    search.advance(buf["flags"], 4)
                else:
                    handle.stage_finish_events(a, 1, f1.data_ptr(), st)
            else:
                handle.stage_finish(a, st)
            if dense:
                record_nodes(buf["flags"].data_ptr())

        def capture():
            """One attempt as a CUDA graph: replaying it costs one launch instead of ~S*(1 + RHS kernels) + 4, which
            is what bounds small and medium ensembles.  A right-hand side that is not capture-safe (host
            synchronisation, CPU work) keeps the eager loop."""
            if not self.use_graphs or callbacks or search is not None:     # callbacks may rebind the system's tensors
                return None                                                # (e.g. ode_sys.dt = ...); the event search looks at the device
            try:
                nfev0, rows0 = self.equ_rhs.nfev, self._rows_used
                g = _capture_graph(dev, lambda: one_attempt(False))
                self.equ_rhs.nfev = nfev0            # the capture pass recorded the calls, it did not execute them
                self._rows_used = rows0
                return g
            except Exception:
                torch.cuda.synchronize(dev)
                return None

        graph, rewarm, capture_pending = None, False, False
        rhs_calls_per_attempt = S + (1 if dense and store.with_f else 0) + (1 if events else 0)
        trace = self.__dict__.get("_attempt_trace")          # optional list: (event, event, working columns) per attempt
        while True:
            if trace is not None:
                ev_pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev_pair[0].record()
            if dense and (self._rows_used + 2) * n > store.capacity:
                # the host knows an upper bound of the row-groups in use (one per attempt) without asking the device;
                # a grown pool has a new address, so the captured attempt is rebuilt
                store.grow(max(2 * store.capacity, (self._rows_used + 2) * n))
                if not first:
                    graph, rewarm = None, True
            if first:
                one_attempt(True)                    # eager: integrate()-entry logic, allocator warm-up
                first = False
                graph = capture()
            elif rewarm:
                # after a compaction (or a grown node pool) the attempt runs on new buffers: one eager attempt first --
                # a right-hand side may meet new shapes, never inside a capture -- and the graph is rebuilt only if a
                # FURTHER attempt turns out to be needed (the compacted tail of an ensemble often finishes in one round:
                # capturing for it costs more than it saves)
                one_attempt(False)
                rewarm, capture_pending = False, True
            elif capture_pending:
                capture_pending = False
                graph = capture()
                if graph is not None:
                    graph.replay()
                    self.equ_rhs.nfev += rhs_calls_per_attempt
                    if dense:
                        self._rows_used += 1
                        store.invalidate()
                else:
                    one_attempt(False)
            elif graph is not None:
                graph.replay()
                self.equ_rhs.nfev += rhs_calls_per_attempt
                if dense:
                    self._rows_used += 1
                    store.invalidate()
            else:
                one_attempt(False)
            self._launches += handle.stage_launches if graph is None else 1
            attempts += 1
            if trace is not None:
                ev_pair[1].record()
                trace.append((ev_pair[0], ev_pair[1], m, graph is not None))
            if callbacks:
                # a single system calls back once per ACCEPTED step like the reference's loop (differential_system.py:
                # 1073-1084); an ensemble once per lock-step round
                fire = True
                if not self.ensemble:
                    acc_now = int(work["accepted"].sum().item())
                    fire, acc_seen = acc_now > acc_seen, acc_now
                if fire:
                    self._fire_callbacks(callbacks)
            if bar is not None:
                bar.update()
            if callbacks or attempts % sync_every == 0:
                active = int(self._counters[1].item())
                if active == 0:
                    break
                if can_compact and m >= 2048 and 2 * active < m:
                    sel = (buf["flags"] & 1).nonzero().squeeze(1)
                    if index is not None:
                        scatter_back()
                    index = sel if index is None else index[sel]
                    m_live = int(sel.numel())
                    # a right-hand side with a tiled kernel (rhs.column_multiple: the linear plugin's dsb_dgemm works on
                    # 128-column tiles) gets a working set padded with INACTIVE copies of its first column, so that the
                    # compacted rounds stay on that kernel instead of falling back to a library GEMM
                    pad = (-m_live) % int(getattr(self.equ_rhs, "column_multiple", 1) or 1)
                    if pad:
                        sel = torch.cat([sel, sel[:1].expand(pad)])
                        index = torch.cat([index, index[:1].expand(pad)])
                    work = {k: v.index_select(v.ndim - 1, index).contiguous() for k, v in full.items()}
                    new_buf = alloc(int(sel.numel()))
                    for k in persistent:
                        new_buf[k] = buf[k].index_select(buf[k].ndim - 1, sel).contiguous()
                    if pad:
                        new_buf["flags"][m_live:] = 0
                        work["status"][m_live:] = cb.TRAJ_DONE
                    buf, m = new_buf, int(sel.numel())
                    consts = {k: (v.index_select(v.ndim - 1, index) if isinstance(v, torch.Tensor) and v.ndim >= 1
                                  and v.shape[-1] == n else v) for k, v in self.__consts.items()}
                    a, y_view, t_view = bind(work, buf, m)
                    self._compactions += 1
                    graph, rewarm = None, True       # new working set, new addresses: one eager attempt, then re-capture
            if attempts > self._step_cap:
                break
        if index is not None:
            scatter_back()
        self._steps_cache = None

    def _run_symplectic(self, tf, callbacks, bar, events=None):
        """Explicit symplectic splitting (fixed dt).  Fused kernel for the tagged N-body right-hand side,
        stage-level kernels around the user's callable otherwise."""
        torch = _torch()
        integ = self.integrator
        n, D, dev = self.n_traj, self.numel, self.device
        stream = cb.current_stream_ptr(dev)
        rhs = self.equ_rhs
        prepare = getattr(rhs, "prepare", None)           # derived device data of the right-hand side (see _run_stagewise)
        if prepare is not None:
            prepare(**self.__consts)
        # drift (position) variables = rows [0, dim0 // 2) of axis 0, kick = the rest (reference integrator_types.py:359-362)
        drift_count = (self.dim[0] // 2) * (D // self.dim[0]) if len(self.dim) else 0
        fused = (self.use_fused and getattr(rhs, "device_plugin", None) == "nbody" and len(self.dim) == 3
                 and self.dim[0] == 2 and self.dim[2] == 3 and not self._want_nodes and not events)
        if fused:
            nb = self.dim[1]
            handle = cb.symp_handle(dev.index or 0, rhs.plugin_id, D, drift_count, integ.tableau_intermediate, strict=self.strict)
            m = torch.as_tensor(self.__consts["m"], dtype=torch.float64, device=dev).contiguous()
            if tuple(m.shape) != (nb,):
                raise ValueError("fused N-body kernel needs masses of shape (%d,) shared by all systems" % nb)
            a = cb.SympRunArgs()
            a.n, a.y, a.t, a.dt = n, self._y.data_ptr(), self._t.data_ptr(), self._dt.data_ptr()
            a.tf = float(tf)
            a.steps, a.status = self.accepted.data_ptr(), self.status.data_ptr()
            a.bodies, a.masses = nb, m.data_ptr()
            a.G, a.eps2 = float(self.__consts.get("G", 1.0)), float(self.__consts.get("eps2", 0.0))
            if not callbacks and bar is None:
                a.max_steps, a.first_call = -1, 1
                handle.run(a, stream)
                self._launches += 1
            else:
                first = True
                while True:
                    a.max_steps, a.first_call = self.callback_every, 1 if first else 0
                    first = False
                    handle.run(a, stream)
                    self._launches += 1
                    self._fire_callbacks(callbacks)
                    if bar is not None:
                        bar.update()
                    if not bool((self.status == cb.TRAJ_RUNNING).any().item()):
                        break
            self._keepalive = m
            return
        handle = cb.symp_handle(dev.index or 0, 100, D, drift_count, integ.tableau_intermediate, strict=self.strict)
        S = integ.stages
        f64 = dict(dtype=torch.float64, device=dev)
        buf = dict(h=torch.zeros(n, **f64), dS=torch.zeros((D, n), **f64), y_stage=torch.empty((D, n), **f64),
                   t_stage=torch.zeros(n, **f64), flags=torch.zeros(n, dtype=torch.int32, device=dev))
        a = cb.SympStageArgs()
        a.n, a.dim = n, D
        a.y, a.t, a.dt = self._y.data_ptr(), self._t.data_ptr(), self._dt.data_ptr()
        for k, v in buf.items():
            setattr(a, k, v.data_ptr())
        a.steps, a.status = self.accepted.data_ptr(), self.status.data_ptr()
        a.tf = float(tf)
        a.counters = self._counters.data_ptr()
        y_view = self._user_shape(buf["y_stage"])
        t_view = self._user_time(buf["t_stage"])
        first, steps = True, 0
        search = None
        if events:
            coef, flags = self._event_tensors(events)
            ev_state = torch.zeros((2 + 2 * cb.MAX_STAGE_EVENTS, n), **f64)
            tf_per = torch.full((n,), float(tf), **f64)
            scratch = torch.zeros(n, dtype=torch.int32, device=dev)
            a.n_events, a.ev_capacity = len(events), self._event_log.capacity
            a.ev_coef, a.ev_flags = coef.data_ptr(), flags.data_ptr()
            a.ev_records, a.ev_count = self._event_log._records.data_ptr(), self._event_log.count.data_ptr()
            a.ev_state, a.tf_per, a.ev_scratch = ev_state.data_ptr(), tf_per.data_ptr(), scratch.data_ptr()
            if self._event_coef_dy is not None:
                a.ev_coef_dy = self._event_coef_dy.data_ptr()
            self._keepalive = (coef, flags, ev_state, tf_per, scratch)   # referenced by raw pointer until the run has finished
            if any(not hasattr(ev, "w") for ev in events):
                from .events import EventSearch
                search = self._event_search = EventSearch(self, events, self.__consts)
                f_start = None
                if search.needs_dy:
                    f_start = self.equ_rhs(self._user_time(self._t), self._user_shape(self._y), **self.__consts).to(torch.float64).reshape(D, n).contiguous()
                search.start(self._t, self._y, f_start)

        store = self._ensure_node_store("rows", need=2) if self._want_nodes else None
        dense = store is not None
        y_committed, t_committed = self._user_shape(self._y), self._user_time(self._t)

        def record_nodes(flags_ptr):
            """Node (t, y[, rhs(t, y)]) of the systems that just stepped (all if flags_ptr is None).  The
            reference's symplectic integrator never refreshes `final_rhs` after the first step
            (integrators/integrator_types.py:385-386), so its later interpolants carry a stale end slope; the nodes
            here hold the slope at their own time."""
            f = None
            if store.with_f:
                f = self.equ_rhs(t_committed, y_committed, **self.__consts).to(torch.float64).reshape(D, n)
                f = f if f.is_contiguous() else f.contiguous()
                keep_f[0] = f
            store.append_rows(self._t, self._y, f, flags_ptr, 4)
            self._rows_used += 1
            self._launches += 2

        keep_f = [None]
        if dense and int(store.node_count.max().item()) == 0:
            record_nodes(None)                                   # node 0 = initial condition

        def one_step(first_call):
            st = cb.current_stream_ptr(dev)
            handle.stage_begin(a, first_call, st)
            for s in range(S):
                f = self.equ_rhs(t_view, y_view, **self.__consts).to(torch.float64).reshape(D, n)
                f = f if f.is_contiguous() else f.contiguous()
                keep[s] = f
                handle.stage_accumulate(a, s, f.data_ptr(), st)
            if events:
                # the first stage's right-hand side is rhs at the step start (dS = 0 there); f1 = rhs after the step
                handle.stage_commit_events(a, 0, None, None, st)
                f1 = self.equ_rhs(t_view, y_view, **self.__consts).to(torch.float64).reshape(D, n)
                keep_f[0] = f1 if f1.is_contiguous() else f1.contiguous()
                if search is not None:
                    roots = search.locate(buf["flags"], 4, 8, ev_state[0], self._t, self._y, buf["y_stage"], keep[0], keep_f[0])
                    handle.stage_commit_events_located(a, keep[0].data_ptr(), keep_f[0].data_ptr(), roots.data_ptr(), st)
                    search.advance(buf["flags"], 4)
                else:
                    handle.stage_commit_events(a, 1, keep[0].data_ptr(), keep_f[0].data_ptr(), st)
            else:
                handle.stage_commit(a, st)
            if dense:
                record_nodes(buf["flags"].data_ptr())

        keep = [None] * S
        graph, regraph = None, False

        def capture_step():
            if not self.use_graphs or callbacks or search is not None:     # same CUDA-graph replay as the explicit-RK stage path
                return None
            try:
                nfev0, rows0 = self.equ_rhs.nfev, self._rows_used
                g = _capture_graph(dev, lambda: one_step(False))
                self.equ_rhs.nfev, self._rows_used = nfev0, rows0
                return g
            except Exception:
                torch.cuda.synchronize(dev)
                return None

        while True:
            if dense and (self._rows_used + 2) * n > store.capacity:
                store.grow(max(2 * store.capacity, (self._rows_used + 2) * n))     # new pool address: rebuild the graph
                if not first:
                    graph, regraph = None, True
            if first:
                one_step(True)
                first = False
                graph = capture_step()
            elif regraph:
                one_step(False)
                regraph = False
                graph = capture_step()
            elif graph is not None:
                graph.replay()
                self.equ_rhs.nfev += S + (1 if (dense and store.with_f) or events else 0)
                if dense:
                    self._rows_used += 1
                    store.invalidate()
            else:
                one_step(False)
            self._launches += (4 + S) if graph is None else 1
            steps += 1
            self._fire_callbacks(callbacks)
            if bar is not None:
                bar.update()
            if callbacks or steps % self.sync_every == 0:
                if int(self._counters[1].item()) == 0:
                    break
            if steps > self._step_cap:
                break
        self._steps_cache = None

    def _run_protocol(self, tf, callbacks, bar):
        """The reference's own loop (differential_system.py:951-1018, 1070-1084) around ANY integrator that implements the
        integrator protocol -- a Richardson wrapper, a user-written `IntegratorTemplate` subclass, or a tableau
        integrator with a custom `adaptation_fn`: final-step clamp, one `integrator(rhs, t, y, constants, timestep)` call
        per accepted step, `y += dState`, `t += dTime`, `dt = new_timestep` unless it was the final step, callbacks.  The
        stepping itself happens in whatever the integrator calls (for this package's classes: the stage-level kernels);
        the loop visits the host once per step, so it serves single systems."""
        torch = _torch()
        if self.ensemble:
            raise NotImplementedError("the protocol loop (Richardson wrappers, custom integrators / adaptation functions) "
                                      "drives one system at a time: ensemble=False")
        integ = self.integrator
        y, t, dt = self._y, self._t, self._dt                      # (D, 1), (1,), (1,) device tensors, advanced in place
        store = self._ensure_node_store("rows", need=2) if self._want_nodes else None
        if store is not None and int(store.node_count.max().item()) == 0:
            f0 = self.equ_rhs(t.reshape(()), self._user_shape(y), **self.__consts).reshape(self.numel, 1).contiguous() if store.with_f else None
            store.append_rows(t, y, f0, None, 4)
            self._rows_used += 1
        tc = float(t[0])
        if abs(tf - tc) < D.epsilon("float64"):
            return
        self._fix_dt_dir(tf, tc)
        if abs(float(dt[0])) > abs(tf - tc):
            dt.fill_(float(np.copysign(abs(tf - tc) * 0.5, tf - tc)))
        steps = 0
        while float(dt[0]) != 0.0 and abs(tf - float(t[0])) >= D.tol_epsilon("float64"):
            tc, h = float(t[0]), float(dt[0])
            final = abs(h + tc) > abs(tf)
            if final:
                h = tf - tc
            new_dt, (dT, dY) = integ(self.equ_rhs, t.reshape(()).clone(), self._user_shape(y).clone(), self.__consts,
                                     timestep=torch.as_tensor(h, dtype=torch.float64, device=self.device))
            y.add_(torch.as_tensor(dY, dtype=torch.float64, device=self.device).reshape(self.numel, 1))
            t.add_(torch.as_tensor(dT, dtype=torch.float64, device=self.device).reshape(1))
            self.accepted.add_(1)
            steps += 1
            if store is not None:
                if (self._rows_used + 2) > store.capacity:
                    store.grow(2 * store.capacity)
                f1 = self.equ_rhs(t.reshape(()), self._user_shape(y), **self.__consts).reshape(self.numel, 1).contiguous() if store.with_f else None
                store.append_rows(t, y, f1, None, 4)
                self._rows_used += 1
            if not final:
                dt.copy_(torch.as_tensor(new_dt, dtype=torch.float64, device=self.device).reshape(1))
            self._fire_callbacks(callbacks)
            if bar is not None:
                bar.update()
            if steps > self._step_cap:
                break
        self.status.zero_()
        self._steps_cache = None
        self._protocol_path = True

    def integrate(self, t=None, callback=None, eta=False, events=None):
        """Integrates every trajectory to time `t` (default: `self.tf`), reference `:912-1101`.

        The whole loop runs on the device without host synchronisation per step.  `callback`
        forces chunked execution: the kernels return every `self.callback_every` accepted steps per
        trajectory and `callback(self)` is invoked in between.  `events`: see `desolver_b200.events`.
        """
        torch = _torch()
        with torch.cuda.device(self.device):       # launches, streams and temporaries all belong to the system's device
            return self._integrate(t, callback, eta, events)

    def _integrate(self, t, callback, eta, events):
        torch = _torch()
        user_events = None
        if events is not None:
            from .events import as_hyperplane, CallableEvent
            user_events = [events] if callable(events) and not isinstance(events, (list, tuple)) else list(events)
            if len(user_events) > cb.MAX_STAGE_EVENTS:
                raise NotImplementedError("at most %d events per integrate() call" % cb.MAX_STAGE_EVENTS)
            # Affine event functions g(t, y[, dy]) = w.y + w_dy.dy + wt*t - c -- HyperplaneEvent objects, or ordinary
            # callables of that form (t - t_mark, y[0] - level, ...: recognised by probing) -- are located in closed form
            # inside the step kernels.  ANY other callable (a radius, a product, per-trajectory constants) is evaluated by
            # the caller's own code between the event-search kernels (events.EventSearch), on the stage-level path.
            try:
                events = [as_hyperplane(ev, self.numel, self.__consts) for ev in user_events]
            except NotImplementedError:
                events = [ev if isinstance(ev, CallableEvent) else CallableEvent(ev) for ev in user_events]
            if not events:
                events = user_events = None
        tf = float(t) if t is not None else self.__tf
        self._detach_history()
        if callback is None:
            callbacks = []
        elif isinstance(callback, (tuple, list)):
            callbacks = list(callback)
        else:
            callbacks = [callback]
        plugin = self._fused_plugin() if self.use_fused else None
        if plugin is not None and self._node_store is not None and self._node_store.kind == "rows":
            plugin = None                  # nodes of this run are row-written: it stays on the stage-level kernels until reset()
        fused_stats = None
        lazy = False
        try:
            if self.host_io and (plugin is None or self.integrator.symplectic):
                raise NotImplementedError("host_io=True needs an explicit Runge-Kutta method with a fused kernel")
            if events and self.host_io:
                raise NotImplementedError("device event detection needs device-resident state (host_io=False)")
            general_events = bool(events) and any(not hasattr(ev, "w") for ev in events)
            if events and plugin is not None and (self.integrator.stages > 7 or self._want_nodes or general_events
                                                  or len(events) > cb.MAX_EVENTS or callbacks or eta):
                # the event-detecting fused kernels stop at 7 stages and write no nodes: a system that records dense
                # output or its step history next to events runs on the stage-level kernels
                plugin = None
            protocol = (not isinstance(self.integrator, (integrators.RungeKuttaIntegrator, integrators.ExplicitSymplecticIntegrator))
                        or getattr(self.integrator, "has_custom_adaptation", False))
            if protocol:
                if events or self.host_io:
                    raise NotImplementedError("integrators driven through the protocol loop (Richardson wrappers, custom "
                                              "integrators / adaptation functions) do not combine with events or host_io")
                plugin = None
            if self.integrator.symplectic or plugin is None:
                self._stats = None
            if protocol:
                bar = None
                if eta:
                    from tqdm.auto import tqdm
                    bar = tqdm(total=None)
                self._run_protocol(tf, callbacks, bar)
                if bar is not None:
                    bar.close()
            elif self.integrator.symplectic:
                bar = None
                if eta:
                    from tqdm.auto import tqdm
                    bar = tqdm(total=None)
                self._run_symplectic(tf, callbacks, bar, events=events)
                if events:
                    self._event_log.events = list(user_events)
                if bar is not None:
                    bar.close()
            elif plugin is None:
                bar = None
                if eta:
                    from tqdm.auto import tqdm
                    bar = tqdm(total=None)
                self._run_stagewise(tf, callbacks, bar, events=events)
                if events:
                    self._event_log.events = list(user_events)
                if bar is not None:
                    bar.close()
            elif self.host_io:
                if callbacks or eta or self.__dense_output or not self.__dict__.get("_pristine"):
                    raise NotImplementedError("host_io=True systems integrate once per reset(), without callbacks, "
                                              "progress bar or dense output")
                fused_stats = self._run_host_pipeline(tf, plugin)
            elif (self._push_target is not None and not callbacks and not eta and not events and not self._want_nodes
                  and self.__dict__.get("_pristine")):
                self._run_push(tf, plugin)
                lazy = True
            elif not callbacks and not eta:
                self._run_fused(tf, plugin, self._step_cap, first_call=True, events=events)
                if events:
                    self._event_log.events = list(user_events)     # StateTuple.event is the caller's own object
                lazy = self.lazy_status and not events
                if not lazy:
                    fused_stats = self._fused_stats()
            else:
                first = True
                bar = None
                if eta:
                    from tqdm.auto import tqdm
                    bar = tqdm(total=None)
                fused_stats = (0, 0, 0, 0)
                while True:
                    self._run_fused(tf, plugin, self.callback_every, first_call=first)
                    first = False
                    st = self._fused_stats()
                    fused_stats = (fused_stats[0] + st[0], fused_stats[1] + st[1], st[2], st[3])
                    self._fire_callbacks(callbacks)
                    if bar is not None:
                        bar.update()
                    if st[3] == 0:
                        break
                if bar is not None:
                    bar.close()
            if self.host_io:
                self._hist_y.append(self._raw_y)              # the pinned result buffer itself (no 25 MB host copy)
                self._hist_t.append(float(tf) if fused_stats[2] + fused_stats[3] == 0 else self._raw_t.cpu())
            else:
                # the newest history row aliases the live state; it is copied only if a later integrate() call is
                # about to advance that state (_detach_history) -- reset() + integrate() loops never pay for it
                self._hist_y.append(self._raw_y)
                self._hist_t.append(self._raw_t)
                self._hist_alias = True
            # error policy: run everything to completion, then raise like the reference does
            # (:1089-1093 wraps integrator_types.py:220-224) if any trajectory failed
            if lazy:
                self._pending = True              # counters are read (and failures raised) on first use
                n_fail = 0
            elif fused_stats is not None:
                prev = self._stats or dict(attempts=0, accepted=0)
                self._stats = dict(attempts=prev["attempts"] + fused_stats[0], accepted=prev["accepted"] + fused_stats[1],
                                   failed=fused_stats[2], running=fused_stats[3])
                n_fail = fused_stats[2]
            else:
                n_fail = int((self.status == cb.TRAJ_TOL_FAIL).sum().item())
            if n_fail:
                cause = etypes.FailedToMeetTolerances(
                    "Failed to integrate {} of {} trajectories to the tolerances required: rtol={}, atol={}".format(
                        n_fail, self.n_traj, self.integrator.rtol, self.integrator.atol))
                raise cause
        except KeyboardInterrupt as e:
            self.__int_status = e
            raise e
        except Exception as e:
            new_e = etypes.FailedIntegration("Failed to integrate system")
            new_e.__cause__ = e
            self.__int_status = new_e
            raise new_e
        else:
            # "Integration terminated upon finding a triggered event." (reference :1055) if a terminal event fired
            terminated = bool(events) and any(ev.is_terminal for ev in events) and \
                bool((self.status == cb.TRAJ_EVENT).any().item())
            if not lazy:
                self.__int_status = 2 if terminated else 1

    # ------------------------------------------------------------------ container protocol
    def __len__(self):
        if self._record_steps:
            return int(self._step_history()[0].shape[0])          # counter + 1, reference :1198-1199
        return len(self._hist_y)

    def __getitem__(self, index):
        torch = _torch()
        if self._record_steps:
            t_rows, y_rows = self._step_history()
            rows = int(t_rows.shape[0])
            if isinstance(index, int):
                if index == -1:                                  # the live state: no gather needed
                    self._materialize()
                    return StateTuple(t=self._user_time(self._raw_t), y=self._user_shape(self._raw_y), event=None)
                if index >= rows or index < -rows:
                    raise IndexError("index {} out of bounds for integrations with {} steps".format(index, rows))
                return StateTuple(t=t_rows[index], y=y_rows[index], event=None)
            if isinstance(index, slice):
                # the reference slices by TIME (:1170-1183): start / stop are times, located by bisection
                lo = self._locate_step(t_rows, index.start) if index.start is not None else 0
                hi = self._locate_step(t_rows, index.stop) + 1 if index.stop is not None else rows
                step = index.step if index.step is not None else 1
                return StateTuple(t=t_rows[lo:hi:step], y=y_rows[lo:hi:step], event=None)
            if self.__dense_output and self._node_store is not None:
                return StateTuple(t=index, y=self.sol(index), event=None)
            # "nearest" stored step exactly as the reference picks it (:1187-1196): the first row at or after the time,
            # unless the row after it is strictly closer
            k = self._locate_step(t_rows, index)
            if k < rows - 1 and not abs(float(t_rows[k]) - float(index)) < abs(float(t_rows[k + 1]) - float(index)):
                k += 1
            return StateTuple(t=t_rows[k], y=y_rows[k], event=None)
        if isinstance(index, int):
            if index >= len(self._hist_y) or index < -len(self._hist_y):
                raise IndexError("index {} out of bounds for integrations with {} steps".format(index, len(self._hist_y)))
            return StateTuple(t=self._user_time(self._hist_time(self._hist_t[index])),
                              y=self._user_shape(self._hist_y[index]), event=None)
        if isinstance(index, slice):
            return StateTuple(t=self.t[index], y=self.y[index], event=None)
        if self.__dense_output and self._node_store is not None:
            return StateTuple(t=index, y=self.sol(index), event=None)
        raise KeyError("time indexing needs dense_output=True in the ensemble engine")

    def __str__(self):
        """Equations, initial conditions, current states, constants and time limits (reference `:1146-1161`)."""
        out = "y({t0}) = {y0}\ndy = {eq}\ny({t}) = {y}\n".format(t0=self.__t0, y0=self[0].y, eq=str(self.equ_rhs),
                                                                 t=self[-1].t, y=self[-1].y)
        if self.constants:
            out += "\nThe constants that have been defined for this system are: \n" + str(self.constants)
        out += "\nThe time limits for this system are:\n"
        out += "t0 = {}, tf = {}, t_current = {}, step_size = {}".format(self.__t0, self.__tf, self[-1].t, self.dt)
        return out

    def _summary_rows(self):
        rows = [("message", self.integration_status), ("nfev", self.nfev), ("njev", self.njev),
                ("sol", self.sol if self.__dense_output else None), ("t0", self.__t0), ("tf", self.__tf)]
        if self.ensemble:
            rows += [("trajectories", self.n_traj), ("method", self.__method.__name__)]
        return rows

    def __repr__(self):
        """The reference's summary (`:1103-1115`); an ensemble prints its size instead of every state."""
        rows = ["{:>10}: {}".format(k, v) for k, v in self._summary_rows()]
        if not self.ensemble:
            rows += ["{:>10}: {}".format("y0", self[0].y), "{:>10}: {}     ".format("Equations", repr(self.equ_rhs)),
                     "{:>10}: {}".format("t", self.t), "{:>10}: {}".format("y", self.y)]
        return "\n".join(rows)

    def _repr_markdown_(self):
        """Notebook rendering (reference `:1117-1144`)."""
        head = "  \n".join("{:>10}: {}".format(k, v) for k, v in self._summary_rows())
        eq = self.equ_rhs._repr_markdown_() if hasattr(self.equ_rhs, "_repr_markdown_") else str(self.equ_rhs)
        tail = "" if self.ensemble else "```  \n{:>10}: {}  \n{:>10}: {}  \n```\n".format("t", self.t, "y", self.y)
        return "```\n{}  \n{:>10}: \n```  \n{}\n \n{}".format(head, "Equations", eq, tail)


def solve_ivp(fun, t_span, y0, method="RK45", t_eval=None, dense_output=False, events=None, vectorized=False, args=None,
              **options):
    """Functional wrapper with the signature of the reference's `solve_ivp`
    (`differential_system.py:1202-1255`, itself modelled on scipy's).  `t_eval` is served by repeated
    `integrate(t)` calls exactly like the reference (`:1245-1248`).  Extra options: `ensemble`, `device`, `strict`.
    Returns a `scipy.optimize.OptimizeResult`-style object with `t`, `y`, `sol`, `nfev`, `status`, `success`,
    `ode_system`; for an ensemble `y` has shape (*dim, N, n_times) and `t` (n_times,) or (N, n_times)."""
    import inspect
    torch = _torch()
    from scipy.optimize import OptimizeResult
    constants = None
    if args is not None:
        fn = fun
        while isinstance(fn, DiffRHS):
            fn = fn.rhs
        names = inspect.getfullargspec(fn)[0][2:]
        constants = {k: v for k, v in zip(names, args)}
    max_step = float(options.get("max_step", np.inf))
    min_step = float(options.get("min_step", 0.0))
    first = min(max(float(options.get("first_step", 1.0)), min_step), max_step)
    # t_eval_mode="steps" (default): repeated integrate(t) like the reference -- every t_eval[k] is an exact stop of the
    # integration; "dense": ONE integration over t_span with the node store, then ONE launch of the T x N Hermite
    # evaluation kernel (scipy's semantics: values at t_eval are interpolated, the steps are those of the free run)
    mode = options.get("t_eval_mode", "steps")
    if mode not in ("steps", "dense"):
        raise ValueError("t_eval_mode must be 'steps' or 'dense'")
    system = OdeSystem(fun, y0, t=t_span, dense_output=dense_output or (t_eval is not None and mode == "dense"), dt=first,
                       atol=options.get("atol"), rtol=options.get("rtol"), constants=constants or {},
                       ensemble=options.get("ensemble", False), device=options.get("device"),
                       strict=options.get("strict", False), dense_capacity=options.get("dense_capacity"))
    system.method = method
    callbacks = list(options.get("callbacks", []))
    if "max_step" in options or "min_step" in options:
        def clip_step(ode_sys):                       # reference :1229-1232
            ode_sys.dt = torch.clamp(ode_sys._dt, min=min_step, max=max_step)
        callbacks.append(clip_step)
    kw = dict(callback=callbacks or None, eta=options.get("show_prog_bar", False), events=events)
    if t_eval is None:
        system.integrate(**kw)
        t_res = system.t
        y_res = torch.movedim(system.y, 0, -1)
    else:
        if torch.is_tensor(t_eval):                   # a torch user passes device tensors; the stop list is host data
            t_eval = t_eval.detach().cpu().numpy()
        t_eval = np.sort(np.asarray(t_eval, dtype=np.float64).reshape(-1))
        t_span = tuple(float(v) for v in t_span)
        if t_eval[0] < min(t_span) or t_eval[-1] > max(t_span):
            raise ValueError(f"Expected `t_eval` to be in the range [{t_span[0]}, {t_span[1]}]")
        if mode == "dense":
            system.integrate(**kw)
            grid = torch.as_tensor(t_eval, dtype=torch.float64, device=system.device)
            if system.ensemble:
                y_res = torch.movedim(system.sol.grid(grid), 0, -1)          # (*dim, N, T)
            else:
                y_res = torch.movedim(system.sol(grid), 0, -1)               # (*dim, T)
            t_res = grid
        elif not callbacks and events is None and not options.get("show_prog_bar", False) and not system._want_nodes:
            # exact stops like the reference's loop of integrate(t) calls; ONE launch where the multi-stop kernel applies
            rows = system.integrate_stops(t_eval)                                 # (T, *dim[, N])
            y_res = torch.movedim(rows, 0, -1)
            grid = torch.as_tensor(t_eval, dtype=torch.float64, device=system.device)
            t_res = grid.expand(system.n_traj, -1) if system.ensemble else grid
        else:
            ts, ys = [], []
            for tq in t_eval:
                system.integrate(t=float(tq), **kw)
                ts.append(system[-1].t.clone())
                ys.append(system[-1].y.clone())
            t_res = torch.stack(ts, dim=-1)
            y_res = torch.stack(ys, dim=-1)
    # like the reference (:1252-1253) both keys carry `ode_system.events`: the list of StateTuple(t, y, event) for a
    # single system, the EventLog for an ensemble
    t_events = y_events = system.events if events is not None else None
    return OptimizeResult(t=t_res, y=y_res, sol=system.sol, t_events=t_events, y_events=y_events, nfev=system.nfev,
                          njev=system.njev, status=system.integration_status, message=system.integration_status,
                          success=system.success, ode_system=system)
