"""Event functions with a device implementation.

The reference detects events after every accepted step by root-finding `event(t, sol(t), **constants)` on the
step's dense-output interpolant (/root/reference/desolver/differential_system.py:29-57, 60-167, 1020-1066), with the
scipy-style protocol: an event is any callable `event(t, y, **constants) -> float` carrying the optional
attributes `is_terminal` (stop at the event) and `direction` (> 0: only rising zero crossings, < 0: only falling,
0: both).

`HyperplaneEvent` is such a callable -- it can be handed to the unmodified reference as it is, which is how
tests/golden/events_*.npz were produced -- whose function is affine,

    g(t, y) = sum_d w[d] * y[d] + wt * t - c,

so that the fused kernels can evaluate it per trajectory in registers, bracket sign changes over each accepted
step and locate the root on the step's cubic Hermite interpolant without leaving the device
(include/desolver_b200.h, dsb_erk_run_args.n_events).  Component crossings (Poincare sections, thresholds,
time marks) are all of this form.
"""
import numpy as np

__all__ = ["HyperplaneEvent", "CallableEvent", "EventSearch", "as_hyperplane", "component_crossing", "component_extremum",
           "EventLog"]


class HyperplaneEvent(object):
    """g(t, y[, dy]) = w . y + w_dy . dy + wt * t - c, with the reference's event attributes `is_terminal`, `direction`
    and -- when `w_dy` is given -- `requires_dstate` (the event then also sees dy = the derivative of the dense-output
    interpolant, /root/reference/desolver/differential_system.py:92-96: extrema of a component, e.g. the successive
    maxima of z of the Lorenz map, are `w_dy = e_z, direction = -1`)."""

    def __init__(self, w, c=0.0, wt=0.0, direction=0, is_terminal=False, name=None, w_dy=None):
        self.w = np.asarray(w, dtype=np.float64).reshape(-1)
        self.w_dy = None
        if w_dy is not None:
            w_dy = np.asarray(w_dy, dtype=np.float64).reshape(-1)
            if w_dy.shape != self.w.shape:
                raise ValueError("w_dy must have as many entries as w")
            if w_dy.any():
                self.w_dy = w_dy
        self.requires_dstate = self.w_dy is not None
        self.c = float(c)
        self.wt = float(wt)
        self.direction = int(np.sign(direction))
        self.is_terminal = bool(is_terminal)
        self.name = name or "hyperplane(w=%s, w_dy=%s, wt=%g, c=%g)" % (
            self.w.tolist(), None if self.w_dy is None else self.w_dy.tolist(), self.wt, self.c)

    def __call__(self, t, y, *args, **constants):
        if self.w_dy is not None:
            dy = args[0]
            return self._affine(t, y) + self._dot(self.w_dy, dy, 0.0)
        return self._affine(t, y)

    def _dot(self, w, y, g):
        n = w.shape[0]
        if y.shape[0] != n:
            lead = 1
            k = 0
            while lead < n:
                lead *= y.shape[k]
                k += 1
            y = y.reshape((n,) + tuple(y.shape[k:]))
        for d in range(n):
            if w[d] != 0.0:
                g = g + w[d] * y[d]
        return g

    def dy_coefficients(self, dim):
        return np.zeros(dim) if self.w_dy is None else self.w_dy

    def _affine(self, t, y, *args, **constants):
        # y is (dim,) for one system, (dim, N) for an ensemble; a multi-axis state (*dim[, N]) is flattened first
        n = self.w.shape[0]
        if y.shape[0] != n:
            lead = 1
            k = 0
            while lead < n:
                lead *= y.shape[k]
                k += 1
            y = y.reshape((n,) + tuple(y.shape[k:]))
        g = self.wt * t - self.c
        for d in range(n):
            if self.w[d] != 0.0:
                g = g + self.w[d] * y[d]
        return g

    def coefficients(self, dim):
        if self.w.shape[0] != dim:
            raise ValueError("event %s has %d weights, the state has %d components" % (self.name, self.w.shape[0], dim))
        return np.concatenate([self.w, [self.wt, self.c]])

    @property
    def flags(self):
        return (1 if self.direction > 0 else (2 if self.direction < 0 else 0)) | (4 if self.is_terminal else 0)

    def __repr__(self):
        return "<HyperplaneEvent %s direction=%d terminal=%s>" % (self.name, self.direction, self.is_terminal)


class CallableEvent(object):
    """ANY event callable of the reference's protocol, `event(t, y[, dy], **constants)` with the optional attributes
    `is_terminal`, `direction`, `requires_dstate` (differential_system.py:29-57).  It is evaluated by the caller's own
    (torch) code, batched over the ensemble: `t` is 0-d and `y` (*dim) for a single system, `t` (N,) and `y` (*dim, N) for
    an ensemble -- and per-trajectory constants arrive exactly as the right-hand side receives them.  The root search
    runs in the event-search kernels (include/desolver_b200.h, dsb_event_search_*) with the callable between them."""

    def __init__(self, fn):
        self.fn = fn
        self.direction = int(np.sign(getattr(fn, "direction", 0)))
        self.is_terminal = bool(getattr(fn, "is_terminal", False))
        self.requires_dstate = bool(getattr(fn, "requires_dstate", False))
        self.name = getattr(fn, "__name__", repr(fn))
        self.source = fn

    def __call__(self, t, y, *args, **constants):
        return self.fn(t, y, *args, **constants)

    def coefficients(self, dim):
        return np.zeros(dim + 2)

    def dy_coefficients(self, dim):
        return np.zeros(dim)

    @property
    def flags(self):
        return (1 if self.direction > 0 else (2 if self.direction < 0 else 0)) | (4 if self.is_terminal else 0)

    @property
    def signature(self):
        return ("callable", id(self.fn), self.flags, self.requires_dstate)

    def __repr__(self):
        return "<CallableEvent %s direction=%d terminal=%s>" % (self.name, self.direction, self.is_terminal)


class EventSearch(object):
    """Per-attempt driver of the event-search kernels for a stage-level run with arbitrary event callables.

    g at the start of a step is carried over from the end of the previous accepted step (`g0`); after an attempt
    `locate()` evaluates every event at the attempt's end state, brackets sign changes over the accepted steps, runs the
    Illinois iteration (`dsb_event_search_step`) with the callables evaluated on the step's cubic Hermite interpolant
    (`dsb_hermite_eval`) in between, and returns the [n_events][n] root times for the event kernel's record / order /
    terminal logic."""

    MAX_ITERATIONS = 64

    def __init__(self, system, events, consts):
        import torch
        from .backend import cuda_backend as cb
        self.sys, self.events, self.consts = system, list(events), consts
        n, E, dev = system.n_traj, len(events), system.device
        self.n, self.E, self.dev = n, E, dev
        f64 = dict(dtype=torch.float64, device=dev)
        self.buf = {k: torch.zeros((E, n), **f64) for k in ("lo", "hi", "flo", "fhi", "x", "tq", "fx", "g0", "g1", "roots")}
        self.side = torch.zeros((E, n), dtype=torch.int32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ev_flags = torch.as_tensor([ev.flags for ev in events], dtype=torch.int32).to(dev)
        self.needs_dy = any(ev.requires_dstate for ev in events)
        self.yq = torch.empty((system.numel, n), **f64)
        self.dq = torch.empty((system.numel, n), **f64) if self.needs_dy else None
        self.iterations = 0
        self.searches = 0
        self._cb = cb

    def _call(self, ev, t_flat, y_flat, dy_flat):
        import torch
        s = self.sys
        t = s._user_time(t_flat)
        y = s._user_shape(y_flat)
        if ev.requires_dstate:
            g = ev(t, y, s._user_shape(dy_flat), **self.consts)
        else:
            g = ev(t, y, **self.consts)
        g = torch.as_tensor(g, dtype=torch.float64, device=self.dev)
        return g.reshape(-1).expand(self.n) if g.numel() == 1 else g.reshape(self.n)

    def evaluate(self, out, t_flat, y_flat, dy_flat):
        for e, ev in enumerate(self.events):
            out[e].copy_(self._call(ev, t_flat, y_flat, dy_flat))

    def start(self, t_flat, y_flat, rhs_flat):
        """g of every event at the current state (the start of the first step)."""
        self.evaluate(self.buf["g0"], t_flat, y_flat, rhs_flat)

    def locate(self, flags, need_mask, skip_mask, t0, t1, p0, p1, m0, m1):
        """After an attempt: p0 / p1 the states before / after it, m0 / m1 the right-hand sides there, t0 / t1 the times;
        `flags` the kernels' flag word (need_mask = accepted, skip_mask = final leg).  Returns the roots tensor."""
        import ctypes as C
        cb, b = self._cb, self.buf
        lib, st = cb.lib(), cb.current_stream_ptr(self.dev)
        self.evaluate(b["g1"], t1, p1, m1)
        s = cb.EventSearchStruct()
        s.n, s.n_events = self.n, self.E
        for k in ("lo", "hi", "flo", "fhi", "x", "tq"):
            setattr(s, k, b[k].data_ptr())
        s.side, s.t0, s.t1, s.counter = self.side.data_ptr(), t0.data_ptr(), t1.data_ptr(), self.counter.data_ptr()
        self.counter.zero_()
        cb.check(lib.dsb_event_search_begin(C.byref(s), b["g0"].data_ptr(), b["g1"].data_ptr(), flags.data_ptr(), int(need_mask),
                                            int(skip_mask), self.ev_flags.data_ptr(), st), "dsb_event_search_begin")
        pending = int(self.counter.item())                    # the one host look per attempt of this mode
        self.searches += pending
        it = 0
        while pending and it < self.MAX_ITERATIONS:
            for e, ev in enumerate(self.events):
                tq = b["tq"][e]
                cb.check(lib.dsb_hermite_eval(t0.data_ptr(), t1.data_ptr(), 1, p0.data_ptr(), p1.data_ptr(), m0.data_ptr(),
                                              m1.data_ptr(), self.n, self.sys.numel, tq.data_ptr(), 1, self.yq.data_ptr(), 0, st),
                         "dsb_hermite_eval")
                if ev.requires_dstate:
                    cb.check(lib.dsb_hermite_eval(t0.data_ptr(), t1.data_ptr(), 1, p0.data_ptr(), p1.data_ptr(), m0.data_ptr(),
                                                  m1.data_ptr(), self.n, self.sys.numel, tq.data_ptr(), 1, self.dq.data_ptr(), 1, st),
                             "dsb_hermite_eval")
                b["fx"][e].copy_(self._call(ev, tq, self.yq, self.dq))
            cb.check(lib.dsb_event_search_step(C.byref(s), b["fx"].data_ptr(), st), "dsb_event_search_step")
            it += 1
            if it % 4 == 0:
                pending = int(self.counter.item())
        self.iterations += it
        cb.check(lib.dsb_event_search_roots(C.byref(s), b["roots"].data_ptr(), st), "dsb_event_search_roots")
        return b["roots"]

    def advance(self, flags, accepted_mask):
        """After the event kernel: g at the end of a step that stays accepted becomes g at the start of the next one."""
        import torch
        keep = (flags & accepted_mask) != 0
        self.buf["g0"].copy_(torch.where(keep.unsqueeze(0), self.buf["g1"], self.buf["g0"]))


def as_hyperplane(event, dim, constants=None):
    """A `HyperplaneEvent` equal to `event` if that callable is an affine function of (t, y) -- which is checked, not
    assumed: the coefficients are read off by probing `event` at the origin and the unit vectors (numpy arrays first,
    CPU torch tensors if the callable insists on torch), and the fit is verified at random points.  Typical scipy-style
    events -- `t - t_mark`, `y[0] - level`, `y[1] + 2*y[0]` -- qualify and then run on the device unchanged; anything
    else (products, per-trajectory constants) raises NotImplementedError.  `requires_dstate` events
    `event(t, y, dy)` are probed in dy as well."""
    if isinstance(event, HyperplaneEvent):
        return event
    with_dy = bool(getattr(event, "requires_dstate", False))
    consts = {k: v for k, v in (constants or {}).items()}
    for v in consts.values():
        if getattr(v, "ndim", 0) >= 1 and type(v).__module__.split(".")[0] == "torch" and v.is_cuda:
            consts = {k: (c.cpu() if hasattr(c, "cpu") else c) for k, c in consts.items()}
            break

    def probe(make):
        def g(t, y, dy=None):
            if with_dy:
                out = event(make(t, scalar=True), make(y), make(np.zeros(dim) if dy is None else dy), **consts)
            else:
                out = event(make(t, scalar=True), make(y), **consts)
            out = np.asarray(out.detach().cpu() if hasattr(out, "detach") else out)
            if out.size != 1:
                raise NotImplementedError("event %r returns %d values for one state: events that depend on per-trajectory "
                                          "constants are not supported on the device" % (event, out.size))
            return float(out.reshape(-1)[0])
        return g
    makers = [lambda v, scalar=False: np.float64(v) if scalar else np.asarray(v, dtype=np.float64)]
    try:
        import torch
        makers.append(lambda v, scalar=False: torch.as_tensor(v, dtype=torch.float64))
    except ImportError:
        pass
    last = None
    for make in makers:
        try:
            g = probe(make)
            zero = np.zeros(dim)
            g0 = g(0.0, zero)
            wt = g(1.0, zero) - g0
            w = np.array([g(0.0, np.eye(dim)[d]) - g0 for d in range(dim)])
            w_dy = np.array([g(0.0, zero, np.eye(dim)[d]) - g0 for d in range(dim)]) if with_dy else None
            rng = np.random.default_rng(12345)
            for _ in range(4):
                tt, yy, dd = float(rng.uniform(-3, 3)), rng.uniform(-3, 3, size=dim), rng.uniform(-3, 3, size=dim)
                want = g(tt, yy, dd)
                fit = float(w @ yy) + wt * tt + g0 + (float(w_dy @ dd) if with_dy else 0.0)
                if not abs(want - fit) <= 1e-9 * (1.0 + abs(want)):
                    raise NotImplementedError("event %r is not an affine function of (t, y): the device detects hyperplane "
                                              "events only (desolver_b200.events.HyperplaneEvent)" % (event,))
            hp = HyperplaneEvent(w, c=-g0, wt=wt, direction=getattr(event, "direction", 0), w_dy=w_dy,
                                 is_terminal=getattr(event, "is_terminal", False), name=getattr(event, "__name__", repr(event)))
            hp.source = event
            return hp
        except NotImplementedError:
            raise
        except Exception as exc:                      # the callable could not digest this array type: try the next one
            last = exc
    raise NotImplementedError("event %r could not be probed for an affine form (%r)" % (event, last))


def component_extremum(index, dim=None, direction=0, is_terminal=False):
    """Event `d y[index] / dt == 0`: an extremum of that component (direction -1: maxima, +1: minima)."""
    if dim is None:
        raise ValueError("component_extremum needs dim= (the number of state components)")
    w = np.zeros(int(dim))
    w[int(index)] = 1.0
    return HyperplaneEvent(np.zeros(int(dim)), w_dy=w, direction=direction, is_terminal=is_terminal,
                           name="dy[%d]/dt == 0" % int(index))


def component_crossing(index, value=0.0, dim=None, direction=0, is_terminal=False):
    """Event `y[index] == value` for a state of `dim` components (flattened index)."""
    if dim is None:
        raise ValueError("component_crossing needs dim= (the number of state components)")
    w = np.zeros(int(dim))
    w[int(index)] = 1.0
    return HyperplaneEvent(w, c=value, direction=direction, is_terminal=is_terminal,
                           name="y[%d] == %g" % (int(index), float(value)))


class EventLog(object):
    """Events found by the device, per trajectory: `count[i]` events (at most `capacity` are kept), event k of
    trajectory i happened at `t[i, k]` in state `y[:, i, k]` and belongs to `events[index[i, k]]`."""

    def __init__(self, events, records, count, dim_shape):
        self.events = list(events)
        self.count = count
        self.capacity = int(records.shape[1])
        self._records = records
        self._dim = tuple(dim_shape)

    @property
    def t(self):
        return self._records[:, :, 0]

    @property
    def y(self):
        import torch
        d = self._records.shape[2] - 2
        return torch.movedim(self._records[:, :, 1:1 + d], 2, 0)

    @property
    def index(self):
        return self._records[:, :, -1].long()

    @property
    def overflowed(self):
        return bool((self.count > self.capacity).any().item())

    def for_trajectory(self, i):
        """The reference's `OdeSystem.events` for trajectory i: a list of StateTuple(t, y, event)."""
        from .differential_system import StateTuple
        k = min(int(self.count[i].item()), self.capacity)
        rec = self._records[i, :k].cpu()
        d = rec.shape[1] - 2
        return [StateTuple(t=rec[j, 0], y=rec[j, 1:1 + d].reshape(self._dim), event=self.events[int(rec[j, -1])])
                for j in range(k)]

    def __len__(self):
        return int(self.count.clamp(max=self.capacity).sum().item())
