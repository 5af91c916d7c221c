"""`CubicHermiteInterp` of the drop-in (reference: /root/reference/desolver/utilities/interpolation.py:7-82).

Same constructor, attributes (`t0 t1 p0 p1 m0 m1 trange tshift`) and call protocol (`interp(t)`, `interp.grad(t)`), but
the data are CUDA tensors -- for an ensemble `p0` is (*dim, N) and `t0`, `t1` are (N,) -- and the evaluation is ONE launch
of `dsb_hermite_eval` (csrc/node_store.cu), which applies the reference's formula in the reference's operation order
with one rounding per operator, per trajectory.
"""
from ..backend import cuda_backend as cb

__all__ = ["CubicHermiteInterp"]


class CubicHermiteInterp(object):
    def __init__(self, t0, t1, p0, p1, m0, m1):
        import torch
        p0 = torch.as_tensor(p0, dtype=torch.float64)
        if not p0.is_cuda:
            p0 = p0.cuda()
        dev = p0.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.p0 = p0.clone()
        self.p1 = torch.as_tensor(p1, **f64).clone()
        self.m0 = torch.as_tensor(m0, **f64).clone()
        self.m1 = torch.as_tensor(m1, **f64).clone()
        self.t0 = torch.as_tensor(t0, **f64).clone()
        self.t1 = torch.as_tensor(t1, **f64).clone()
        # ensemble <=> one (t0, t1) pair per trajectory on the trailing axis of the state
        self._n = int(self.t0.shape[0]) if self.t0.ndim == 1 else 1
        if self.t0.ndim == 1 and (self.p0.ndim == 0 or self.p0.shape[-1] != self._n):
            raise ValueError("t0 of shape (N,) needs states of shape (*dim, N)")
        self._dim = self.p0.numel() // self._n

    @property
    def trange(self):
        return self.t1 - self.t0

    @property
    def tshift(self):
        return self.t0

    def _eval(self, t_eval, derivative):
        import torch
        dev = self.p0.device
        tq = torch.as_tensor(t_eval, dtype=torch.float64, device=dev)
        if tq.ndim == 0:
            tq, sq = tq.reshape(1), 0
        elif tuple(tq.shape) == (self._n,):
            tq, sq = tq.contiguous(), 1
        else:
            raise ValueError("t_eval must be a scalar or hold one time per trajectory")
        out = torch.empty_like(self.p0)
        flat = lambda x: x.reshape(self._dim, self._n)                 # noqa: E731
        with torch.cuda.device(dev):
            cb.check(cb.lib().dsb_hermite_eval(self.t0.reshape(-1).data_ptr(), self.t1.reshape(-1).data_ptr(),
                                               1 if self.t0.ndim == 1 else 0, flat(self.p0).data_ptr(), flat(self.p1).data_ptr(),
                                               flat(self.m0).data_ptr(), flat(self.m1).data_ptr(), self._n, self._dim,
                                               tq.data_ptr(), sq, out.data_ptr(), 1 if derivative else 0,
                                               cb.current_stream_ptr(dev)), "dsb_hermite_eval")
        return out

    def __call__(self, t_eval):
        return self._eval(t_eval, False)

    def grad(self, t_eval):
        return self._eval(t_eval, True)

    def __repr__(self):
        return "<CubicHermiteInterp(t0={}, t1={}, shape={})>".format(self.t0.tolist(), self.t1.tolist(), tuple(self.p0.shape))
