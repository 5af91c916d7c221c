"""Integrator classes of the drop-in: same names, class attributes and protocol as the reference's
(/root/reference/desolver/integrators/integrator_template.py:10-93, integrator_types.py:19-160, 326-417).

Inside `OdeSystem.integrate()` an instance is only a typed description (tableau, order, tolerances, device) from which
the system builds a `dsb_erk_handle` / `dsb_symp_handle`: the kernels of libdesolver_b200.so do the stepping, the retry
chain and the step-size control.  The reference's integrator PROTOCOL is nevertheless complete, so that the reference's
own `OdeSystem` (or any caller written against it) can drive these classes:

    integrator(rhs, initial_time, initial_state, constants, timestep) -> (new_timestep, (dTime, dState))
    integrator.dense_output() -> (t_end, CubicHermiteInterp)
    integrator.update_timestep() / integrator.adaptation_fn            (custom step-size controllers)
    .dState .dTime .order .is_adaptive .is_implicit .is_fsal .symplectic .solver_dict .staggered_mask

One protocol step runs on the stage-level kernels with `rhs` evaluated between them on the same stream.
"""
import abc

import numpy as np

from .. import backend as D

__all__ = ["IntegratorTemplate", "TableauIntegrator", "RungeKuttaIntegrator", "ExplicitSymplecticIntegrator"]


def _torch():
    import torch
    return torch


class IntegratorTemplate(abc.ABC):
    symplectic = False
    order = None

    def __init__(self):
        self.solver_dict = None
        self.__custom_adaptation_fn = None

    @property
    def is_adaptive(self):
        return False

    # ---- custom step-size controllers (reference integrator_template.py:22-36) ----
    @property
    def adaptation_fn(self):
        return self.__custom_adaptation_fn if self.__custom_adaptation_fn else self.update_timestep

    @adaptation_fn.setter
    def adaptation_fn(self, update_step_fn):
        self._seed_solver_dict()
        try:
            outputs = update_step_fn(self)
            assert isinstance(outputs, tuple)
            assert len(outputs) == 2
            _ = 2.0 * outputs[0]
            assert isinstance(outputs[1], bool)
        except AssertionError:
            raise ValueError("Step adaptation function must be a function that takes `self` and returns a tuple of exactly "
                             "two outputs of the form: (timestep[float], redo_step[bool])")
        self.__custom_adaptation_fn = update_step_fn

    def _seed_solver_dict(self):
        """The setter validates a controller by CALLING it (reference :27-35), so `solver_dict` must hold the entries a
        controller reads; the reference's constructor fills them with zeros (integrator_types.py:143-149)."""
        torch = _torch()
        if self.solver_dict is None or "initial_state" in self.solver_dict:
            return
        dev = self.device if self.device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        z = torch.zeros(tuple(getattr(self, "dim", ())), dtype=torch.float64, device=dev)
        self.solver_dict.update(initial_state=z, diff=z.clone(), dState=z.clone(),
                                timestep=torch.ones((), dtype=torch.float64, device=dev))

    @property
    def has_custom_adaptation(self):
        return self.__custom_adaptation_fn is not None

    def update_timestep(self, ignore_custom_adaptation=False):
        """The reference's controller over `solver_dict` (integrator_template.py:46-93), for ONE system (one global L2
        norm): `(new_timestep, redo_step)`.  The kernels carry their own fused copy of this arithmetic per trajectory;
        this method exists for callers that inspect the controller or chain it inside a custom `adaptation_fn`."""
        torch = _torch()
        sd = self.solver_dict
        if self.has_custom_adaptation and not ignore_custom_adaptation:
            return self.adaptation_fn(self)
        with torch.no_grad():
            initial_state, diff, timestep = sd["initial_state"], sd["diff"], sd["timestep"]
            safety_factor, atol, rtol, dState, order = sd["safety_factor"], sd["atol"], sd["rtol"], sd["dState"], sd["order"]
            scaling = torch.maximum(torch.abs(initial_state), torch.abs(dState / timestep))
            if "system_scaling" in sd:
                scaling = 0.8 * sd["system_scaling"] + 0.2 * scaling
            sd["system_scaling"] = scaling
            tol = atol + rtol * scaling
            epsilon_current = torch.reciprocal(torch.linalg.norm((diff / tol).reshape(-1)))
            epsilon_last, epsilon_last_last = sd.get("epsilon_last"), sd.get("epsilon_last_last")
            one = torch.ones_like(epsilon_current)

            def guarded(eps, expo):
                k = eps ** expo
                return torch.where(k > 0.0, k, one)
            if epsilon_last is None:
                corr = torch.where(epsilon_current > 0.0, epsilon_current ** (1.0 / order), one)
                sd["epsilon_last"] = epsilon_current
            elif epsilon_last_last is None:
                corr = torch.where(epsilon_current > 0.0, epsilon_current ** (1.0 / order), one)
                corr = corr * torch.where(epsilon_last > 0.0, epsilon_last ** (1.0 / order), one)
                sd["epsilon_last_last"], sd["epsilon_last"] = epsilon_last, epsilon_current
            else:
                k1, k2, k3 = sd.get("__adapt_k1", 0.55), sd.get("__adapt_k2", -0.27), sd.get("__adapt_k3", 0.05)
                corr = guarded(epsilon_current, k1 / order) * guarded(epsilon_last, k2 / order) * guarded(epsilon_last_last, k3 / order)
                sd["epsilon_last_last"], sd["epsilon_last"] = epsilon_last, epsilon_current
            corr = 1 + torch.arctan(safety_factor * corr - 1)
            return corr * timestep, bool(corr < 0.9 ** 2)

    @abc.abstractmethod
    def __call__(self, rhs, initial_time, initial_state, constants, timestep):
        pass

    @abc.abstractmethod
    def dense_output(self):
        pass

    @classmethod
    def __str__(cls):
        return cls.__name__

    def __repr__(self):
        return "<{}({},{},{},{},{})>".format(self.__class__.__name__, self.dim, self.dtype, self.rtol, self.atol,
                                             self.device)


class TableauIntegrator(IntegratorTemplate, abc.ABC):
    tableau_intermediate = None
    __order__ = None

    def __init__(self, sys_dim, dtype, rtol=None, atol=None, device=None):
        super().__init__()
        self.dim = tuple(int(i) for i in sys_dim)
        self.numel = int(np.prod(self.dim)) if len(self.dim) else 1
        self.dtype = dtype
        self.device = device
        # defaults: 32 * D.epsilon(dtype) = 128 eps (reference integrator_types.py:37-38).  Like the reference, a
        # tolerance may be an array of the state's shape (one value per component); those are kept as float64 arrays.
        self.rtol = self._tolerance(rtol, dtype)
        self.atol = self._tolerance(atol, dtype)
        self.dState = None
        self.dTime = None
        self.initial_state = None
        self.initial_time = None
        self.initial_rhs = None
        self.final_rhs = None
        self._explicit = True
        self._adaptive = False
        self._fsal = False
        self._last_call = None            # (rhs, constants) of the last protocol step: dense_output() may need rhs(t1, y1)

    def _tolerance(self, value, dtype):
        if value is None:
            return 32 * D.epsilon(dtype if dtype is not None else "float64")
        if hasattr(value, "detach"):
            value = value.detach().cpu().numpy()
        arr = np.asarray(value, dtype=np.float64)
        if arr.size == 1:
            return float(arr.reshape(-1)[0])
        return np.ascontiguousarray(np.broadcast_to(arr, self.dim))

    @property
    def has_vector_tolerances(self):
        return isinstance(self.rtol, np.ndarray) or isinstance(self.atol, np.ndarray)

    def tolerance_arrays(self, device):
        """(rtol, atol) as flat float64 device tensors of `numel` entries, for the kernels' per-component tolerance
        arguments (dsb_erk_stage_args.rtol_vec / atol_vec)."""
        torch = _torch()
        out = []
        for v in (self.rtol, self.atol):
            arr = np.broadcast_to(np.asarray(v, dtype=np.float64), self.dim).reshape(-1) if len(self.dim) else np.asarray([v])
            out.append(torch.as_tensor(np.array(arr, dtype=np.float64), dtype=torch.float64).to(device))
        return out

    @classmethod
    def integrator_order(cls):
        return cls.__order__

    @property
    def order(self):
        return self.__class__.__order__

    @property
    def is_implicit(self):
        return not self._explicit

    @property
    def is_explicit(self):
        return self._explicit

    @property
    def is_fsal(self):
        return self._fsal

    @property
    def is_adaptive(self):
        return self._adaptive

    @property
    def stages(self):
        return self.tableau_intermediate.shape[0]

    def _prepare(self, initial_time, initial_state, timestep):
        """Common entry of a protocol step: device copies in the library's SoA layout."""
        torch = _torch()
        y0 = torch.as_tensor(initial_state, dtype=torch.float64)
        if not y0.is_cuda:
            y0 = y0.to(self.device if self.device is not None else "cuda")
        dev = y0.device
        ensemble = y0.ndim == len(self.dim) + 1
        n = int(y0.shape[-1]) if ensemble else 1
        f64 = dict(dtype=torch.float64, device=dev)
        y = y0.reshape(self.numel, n).contiguous().clone()
        t = torch.as_tensor(initial_time, **f64).expand(n).contiguous().clone()
        dt = torch.as_tensor(timestep, **f64).expand(n).contiguous().clone()
        shape = tuple(self.dim) + ((n,) if ensemble else ())
        return y0, y, t, dt, dev, ensemble, n, shape

    def dense_output(self):
        """`(t_end, CubicHermiteInterp)` of the last protocol step (reference integrator_types.py:109-119): the cubic
        Hermite interpolant through (t0, y0, rhs(t0, y0)) and (t1, y1, rhs(t1, y1))."""
        from ..utilities.interpolation import CubicHermiteInterp
        if self.dState is None or self.initial_rhs is None:
            raise ValueError("dense_output() needs a completed step: call the integrator first")
        t1 = self.initial_time + self.dTime
        y1 = self.initial_state + self.dState
        if self.final_rhs is None:
            rhs, constants = self._last_call
            self.final_rhs = rhs(t1, y1, **constants)
        return t1, CubicHermiteInterp(self.initial_time, t1, self.initial_state, y1, self.initial_rhs, self.final_rhs)


class RungeKuttaIntegrator(TableauIntegrator, abc.ABC):
    tableau_final = None

    def __init__(self, sys_dim, dtype, rtol=None, atol=None, device=None):
        super().__init__(sys_dim, dtype, rtol=rtol, atol=atol, device=device)
        ti, tf = self.__class__.tableau_intermediate, self.__class__.tableau_final
        self._adaptive = tf.shape[0] == 2                                       # reference :133
        self._fsal = bool(np.all(ti[-1, 1:] == tf[0, 1:]))                      # reference :134
        self._explicit = all((ti[c, c + 1:] == 0.0).all() for c in range(ti.shape[0]))
        self.solver_dict = dict(safety_factor=0.8, order=self.order, atol=self.atol, rtol=self.rtol,
                                redo_count=0, num_step_retries=64)
        self.solver_dict_keep_keys = set(self.solver_dict.keys())

    def __call__(self, rhs, initial_time, initial_state, constants, timestep):
        """The reference's integrator protocol (`integrators/integrator_template.py:38-40`, called from
        `differential_system.py:1003-1004`): ONE accepted step, with the retry chain of
        `integrator_types.py:162-228`, for every trajectory of `initial_state` (*dim[, N]).

        Returns `(new_timestep, (dTime, dState))`; times and step sizes are 0-d tensors for a single system and
        (N,) tensors for an ensemble (`initial_state.ndim == len(dim) + 1`).  Runs on the stage-level kernels of
        libdesolver_b200.so with the callable `rhs` evaluated between them; raises `FailedToMeetTolerances` if
        any trajectory exhausts its 64 retries.  With a custom `adaptation_fn` the accept / retry decision is the
        caller's function, evaluated on the host after every attempt."""
        torch = _torch()
        from .. import exception_types as etypes
        from ..backend import cuda_backend as cb
        y0, y, t, dt, dev, ensemble, n, shape = self._prepare(initial_time, initial_state, timestep)
        D_ = self.numel
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.solver_dict = {k: v for k, v in self.solver_dict.items() if k in self.solver_dict_keep_keys}
        self._last_call = (rhs, constants)
        with torch.cuda.device(dev):
            handle = cb.erk_handle(dev.index or 0, 100, D_, self.tableau_intermediate, self.tableau_final, self.order)
            S = self.stages
            chunks = max(1, min(4096, (D_ + 127) // 128))
            buf = dict(h=torch.zeros(n, **f64), h0=torch.zeros(n, **f64), y_stage=torch.empty((D_, n), **f64),
                       t_stage=torch.zeros(n, **f64), dY=torch.zeros((D_, n), **f64), scale=torch.zeros((D_, n), **f64),
                       hist=torch.ones((3, n), **f64), partial=torch.zeros((chunks, n), **f64),
                       trial=torch.zeros(n, **i32), flags=torch.zeros(n, **i32))
            status = torch.zeros(n, **i32)
            counters = torch.zeros(8, dtype=torch.int64, device=dev)
            a = cb.ErkStageArgs()
            a.n, a.dim, a.chunks = n, D_, chunks
            a.y, a.t, a.dt = y.data_ptr(), t.data_ptr(), dt.data_ptr()
            for k, v in buf.items():
                setattr(a, k, v.data_ptr())
            # tf = +/-inf in the direction of the step: the final-step clamp belongs to the caller's loop (:997-1002)
            direction = float(torch.sign(dt[0]).item()) or 1.0
            a.tf = direction * float("inf")
            if self.has_vector_tolerances:
                tol_keep = self.tolerance_arrays(dev)
                a.rtol_vec, a.atol_vec = tol_keep[0].data_ptr(), tol_keep[1].data_ptr()
            else:
                a.rtol, a.atol = float(self.rtol), float(self.atol)
            a.status, a.counters, a.one_step = status.data_ptr(), counters.data_ptr(), 1
            stream = cb.current_stream_ptr(dev)
            y_view = buf["y_stage"].reshape(shape)
            t_view = buf["t_stage"] if ensemble else buf["t_stage"].reshape(())
            keep = [None] * S

            def stages():
                for s in range(S):
                    handle.stage_state(a, s, stream)
                    k = rhs(t_view, y_view, **constants).to(torch.float64).reshape(D_, n)
                    keep[s] = k if k.is_contiguous() else k.contiguous()
                    a.K[s] = keep[s].data_ptr()

            if self.has_custom_adaptation and self.is_adaptive:
                if n != 1:
                    raise NotImplementedError("a custom adaptation_fn decides for one system at a time (ensemble=False)")
                self._custom_step(handle, a, stages, keep, y0, y, t, dt, buf, shape, stream)
            else:
                first = True
                for _ in range(66):
                    handle.stage_begin(a, first, stream)
                    first = False
                    stages()
                    handle.stage_finish(a, stream)
                    if int(counters[1].item()) == 0:
                        break
                if bool((status == cb.TRAJ_TOL_FAIL).any().item()):
                    raise etypes.FailedToMeetTolerances(
                        "Failed to integrate system to the tolerances required: rtol={}, atol={}".format(self.rtol, self.atol))
            t0 = torch.as_tensor(initial_time, **f64).expand(n)
            self.initial_state = y0.clone()
            self.initial_time = (t0 if ensemble else t0.reshape(())).clone()
            self.initial_rhs = keep[0].reshape(shape).clone()                    # K_0 = rhs(t0, y0) on every attempt
            # FSAL: the last stage is evaluated at (t0 + h, y0 + dState) (reference :304-306); else on demand
            self.final_rhs = keep[S - 1].reshape(shape).clone() if self.is_fsal else None
            self.dState = (y - y0.reshape(D_, n)).reshape(shape)
            self.dTime = (t - t0) if ensemble else (t - t0).reshape(())
            new_dt = dt if ensemble else dt.reshape(())
        return new_dt, (self.dTime, self.dState)

    def _custom_step(self, handle, a, stages, keep, y0, y, t, dt, buf, shape, stream):
        """Reference `__call__` (integrator_types.py:162-228) with the step-size decision left to `adaptation_fn`:
        stage states on the device, increment / error estimate / decision on the host side of the seam."""
        torch = _torch()
        from .. import exception_types as etypes
        tab = torch.as_tensor(self.tableau_final, dtype=torch.float64, device=y.device)
        b, db = tab[0, 1:], tab[0, 1:] - tab[1, 1:]
        current = dt.clone()
        timestep = dt.clone()
        sd = self.solver_dict
        retries = int(sd.get("num_step_retries", 64))
        for attempt in range(retries + 1):
            if attempt:
                sd["redo_count"] = attempt
                timestep = torch.minimum(timestep, current) if float(current[0]) >= 0 else -torch.minimum(timestep.abs(), current.abs())
            dt.copy_(timestep)
            buf["trial"].zero_()
            buf["flags"].zero_()
            handle.stage_begin(a, True, stream)               # h = dt (tf = inf: no clamp)
            stages()
            K = torch.stack([k.reshape(shape) for k in keep], dim=-1)
            h = timestep.reshape(())
            dState = h * torch.sum(K * b, dim=-1)
            sd.update(diff=h * torch.sum(db * K, dim=-1), initial_state=y0, initial_time=t.reshape(()).clone(),
                      timestep=h.clone(), atol=self.atol, rtol=self.rtol, dState=dState)
            new_t, redo = self.adaptation_fn(self)
            timestep = torch.as_tensor(new_t, dtype=torch.float64, device=y.device).reshape(1)
            if not redo:
                y.add_(dState.reshape(y.shape))
                t.add_(h)
                dt.copy_(timestep)
                return
        raise etypes.FailedToMeetTolerances(
            "Failed to integrate system to the tolerances required: rtol={}, atol={}".format(self.rtol, self.atol))


class ExplicitSymplecticIntegrator(TableauIntegrator):
    symplectic = True

    def __init__(self, sys_dim, dtype=None, staggered_mask=None, rtol=None, atol=None, device=None):
        super().__init__(sys_dim, dtype, rtol=rtol, atol=atol, device=device)
        if staggered_mask is not None:
            # unreachable through the reference's OdeSystem and crashes in its constructor
            # (SURVEY.md fact 7); only the default kick mask is live.
            raise NotImplementedError("custom staggered_mask is not supported (it is dead code in the reference)")
        # kick (momentum) variables = second half of axis 0 (reference :359-362)
        self.staggered_mask = np.zeros(self.dim, dtype=bool)
        self.staggered_mask[self.dim[0] // 2:] = True
        self.solver_dict = dict()

    def __call__(self, rhs, initial_time, initial_state, constants, timestep):
        """One fixed step of the splitting scheme for every system of `initial_state` (reference
        integrator_types.py:375-417) on the stage-level kernels: `(timestep, (dTime, dState))`."""
        torch = _torch()
        from ..backend import cuda_backend as cb
        y0, y, t, dt, dev, ensemble, n, shape = self._prepare(initial_time, initial_state, timestep)
        D_ = self.numel
        f64 = dict(dtype=torch.float64, device=dev)
        self._last_call = (rhs, constants)
        drift_count = (self.dim[0] // 2) * (D_ // self.dim[0]) if len(self.dim) else 0
        with torch.cuda.device(dev):
            handle = cb.symp_handle(dev.index or 0, 100, D_, drift_count, self.tableau_intermediate)
            buf = dict(h=torch.zeros(n, **f64), dS=torch.zeros((D_, n), **f64), y_stage=torch.empty((D_, n), **f64),
                       t_stage=torch.zeros(n, **f64), flags=torch.zeros(n, dtype=torch.int32, device=dev))
            steps = torch.zeros(n, dtype=torch.int32, device=dev)
            status = torch.zeros(n, dtype=torch.int32, device=dev)
            counters = torch.zeros(8, dtype=torch.int64, device=dev)
            a = cb.SympStageArgs()
            a.n, a.dim = n, D_
            a.y, a.t, a.dt = y.data_ptr(), t.data_ptr(), dt.data_ptr()
            for k, v in buf.items():
                setattr(a, k, v.data_ptr())
            a.steps, a.status, a.counters = steps.data_ptr(), status.data_ptr(), counters.data_ptr()
            direction = float(torch.sign(dt[0]).item()) or 1.0
            a.tf = direction * float("inf")
            stream = cb.current_stream_ptr(dev)
            y_view = buf["y_stage"].reshape(shape)
            t_view = buf["t_stage"] if ensemble else buf["t_stage"].reshape(())
            handle.stage_begin(a, True, stream)
            f0 = None
            keep = []
            for s in range(self.stages):
                f = rhs(t_view, y_view, **constants).to(torch.float64).reshape(D_, n)
                f = f if f.is_contiguous() else f.contiguous()
                keep.append(f)
                if s == 0:
                    f0 = f.reshape(shape).clone()
                handle.stage_accumulate(a, s, f.data_ptr(), stream)
            handle.stage_commit(a, stream)
            t0 = torch.as_tensor(initial_time, **f64).expand(n)
            self.initial_state = y0.clone()
            self.initial_time = (t0 if ensemble else t0.reshape(())).clone()
            self.initial_rhs = f0
            self.final_rhs = None
            self.dState = (y - y0.reshape(D_, n)).reshape(shape)
            self.dTime = (t - t0) if ensemble else (t - t0).reshape(())
            step = torch.as_tensor(timestep, **f64)
        return step, (self.dTime, self.dState)
