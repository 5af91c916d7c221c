"""A right-hand side of the CALLER inside the persistent fused kernel: `rhs_plugins.register_device_rhs` compiles the CUDA
body with the kernel templates of the library into a plug-in shared object, registers it (dsb_rhs_register) and tags the
Python callable.  Example: the Roessler system of `desolver_b200.benchmarks` (prebuilt by `__graft_entry__.build()`).

Oracle: the SAME callable, untagged, through the stage-level kernels (which the golden fixtures pin against the
reference for arbitrary callables), and -- where baseline/_ref travelled -- the unmodified reference itself."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _untagged(t, y, a=0.2, b=0.2, c=5.7):
    return torch.stack([-y[1] - y[2], y[0] + a * y[1], b + y[2] * (y[0] - c)])


def _ensemble(n, seed=3):
    rng = np.random.default_rng(seed)
    return rng.uniform([-8, -8, 0], [8, 8, 12], (n, 3)).T.copy()


def test_user_plugin_runs_in_the_fused_kernel_and_equals_the_callable_path():
    de = _de()
    rossler = de.benchmarks.rossler_plugin()
    assert rossler.plugin_id >= 1000 and rossler.device_plugin in de.rhs_plugins.FUSED
    n = 20000
    y0 = torch.as_tensor(_ensemble(n))
    c = torch.as_tensor(np.random.default_rng(1).uniform(4.0, 7.0, n)).cuda()         # one parameter per trajectory
    kw = dict(t=(0.0, 6.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a = de.OdeSystem(rossler, y0, constants=dict(a=0.2, b=0.2, c=c), **kw)
    a.method = "RK45"
    assert a._fused_plugin() is not None
    a.integrate()
    launches = a._launches
    b = de.OdeSystem(_untagged, y0, constants=dict(a=0.2, b=0.2, c=c), **kw)
    b.method = "RK45"
    b.integrate()
    assert launches <= 2 < b._launches                              # ONE persistent launch against one per lock-step round
    same = (a.accepted == b.accepted) & (a.rejected == b.rejected)
    err = ((a[-1].y - b[-1].y).abs().amax(0) / b[-1].y.abs().amax(0).clamp_min(1.0)).max()
    print("user plugin: counts identical %.5f, max rel difference %.3g, %d attempts" % (float(same.double().mean()), float(err),
                                                                                      int((a.accepted + a.rejected).sum())))
    assert float(same.double().mean()) >= 0.999 and float(err) <= 1e-10 and a.success
    assert a.nfev == b.nfev
    # bitwise reproducible, and every other fused tableau instantiation is there too
    a2 = de.OdeSystem(rossler, y0, constants=dict(a=0.2, b=0.2, c=c), **kw)
    a2.method = "RK45"
    a2.integrate()
    assert torch.equal(a2[-1].y, a[-1].y) and torch.equal(a2.accepted, a.accepted)
    for method in ("RK87", "DOPRI45", "RK4"):
        u = de.OdeSystem(rossler, y0[:, :512], constants=dict(a=0.2, b=0.2, c=c[:512]), **dict(kw, dt=5e-3))
        v = de.OdeSystem(_untagged, y0[:, :512], constants=dict(a=0.2, b=0.2, c=c[:512]), **dict(kw, dt=5e-3))
        for s in (u, v):
            s.method = method
            s.integrate()
        assert u._launches <= 2
        # (RK8(7) from dt0 = 5e-3: the first error estimates are rounding noise, so a few step counts ride on the last ulp
        # of the two code paths' different FMA contractions -- SURVEY.md fact 6)
        frac = float((u.accepted == v.accepted).double().mean())
        assert frac == 1.0 if method != "RK87" else frac >= 0.95, (method, frac)
        assert float(((u[-1].y - v[-1].y).abs().amax(0) / v[-1].y.abs().amax(0).clamp_min(1.0)).max()) <= 1e-8


def test_user_plugin_with_dense_output_events_and_single_systems():
    de = _de()
    rossler = de.benchmarks.rossler_plugin()
    y0 = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    kw = dict(t=(0.0, 12.0), dt=1e-2, rtol=1e-9, atol=1e-9, dense_output=True)

    def section(t, y, **k):
        return y[0]
    section.direction = 1
    a = de.OdeSystem(rossler, y0, **kw)
    b = de.OdeSystem(_untagged, y0, **kw)
    for s in (a, b):
        s.method = "RK45"
        s.integrate(events=[section])
    assert len(a) == len(b) and float((a.t - b.t).abs().max()) <= 1e-9 and float((a.y - b.y).abs().max()) <= 1e-8
    q = torch.linspace(0.1, 11.9, 7, dtype=torch.float64, device="cuda")
    assert float((a.sol(q) - b.sol(q)).abs().max()) <= 1e-8
    assert len(a.events) == len(b.events) >= 1
    assert max(abs(float(e.t) - float(f.t)) for e, f in zip(a.events, b.events)) <= 1e-9
    # STRICT arithmetic: the plug-in carries the fast flavour only -> stage-level kernels, same callable
    s = de.OdeSystem(rossler, y0, strict=True, **kw)
    s.method = "RK45"
    assert s._fused_plugin() is None
    s.integrate()
    assert len(s) == len(b)


def test_time_dependent_user_plugin_with_transcendental_functions():
    """TIME = true plug-in (stage time t + c_s h), sin / cos in the body, per-trajectory and shared parameters, two ids
    registered side by side."""
    de = _de()
    pend = de.benchmarks.driven_pendulum_plugin()
    ross = de.benchmarks.rossler_plugin()
    assert pend.plugin_id != ross.plugin_id and pend.plugin_id >= 1000
    n = 4096
    rng = np.random.default_rng(2)
    y0 = torch.as_tensor(rng.uniform([-3, -2], [3, 2], (n, 2)).T.copy())
    f = torch.as_tensor(rng.uniform(0.5, 1.5, n)).cuda()
    kw = dict(t=(1.0, 9.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, constants=dict(q=0.5, f=f, w=2.0 / 3.0))

    def untagged(t, y, q, f, w):
        return torch.stack([y[1], -torch.sin(y[0]) - q * y[1] + f * torch.cos(w * t)])
    a = de.OdeSystem(pend, y0, **kw)
    b = de.OdeSystem(untagged, y0, **kw)
    for s in (a, b):
        s.method = "RK45"
        s.integrate()
    assert a._launches <= 2 < b._launches
    same = float(((a.accepted == b.accepted) & (a.rejected == b.rejected)).double().mean())
    err = float(((a[-1].y - b[-1].y).abs().amax(0) / b[-1].y.abs().amax(0).clamp_min(1.0)).max())
    print("driven pendulum plug-in: counts identical %.5f, max rel difference %.3g" % (same, err))
    assert same >= 0.999 and err <= 1e-9                       # (device sin / cos against torch's: last-ulp differences)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "desolver")), reason="baseline/_ref not present")
def test_user_plugin_against_the_live_reference():
    de = _de()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import differential_vs_reference as dv
    ref = dv._import_reference()
    rossler = de.benchmarks.rossler_plugin()
    n = 24
    y0 = _ensemble(n, seed=9)
    kw = dict(t=(0.0, 5.0), dt=1e-2, rtol=1e-9, atol=1e-9)
    a = de.OdeSystem(rossler, torch.as_tensor(y0), ensemble=True, **kw)
    a.method = "RK45"
    a.integrate()
    acc, y = a.accepted.cpu().numpy(), a[-1].y.cpu().numpy()
    for i in range(n):
        r = ref.OdeSystem(de.benchmarks.rossler, y0=y0[:, i].copy(), **kw)         # the same Python callable, numpy arrays
        r.method = "RK45"
        r.integrate()
        assert len(r) - 1 == acc[i]
        assert np.max(np.abs(y[:, i] - np.asarray(r.y[-1]))) <= 1e-10 * max(1.0, np.max(np.abs(r.y[-1])))


def test_traced_plugin_needs_no_cuda_source():
    """`register_device_rhs(fn, dim=...)` without a body: the callable is evaluated once on symbolic scalars and the
    recorded expressions become the kernel's right-hand side (rhs_plugins.trace_body)."""
    de = _de()
    body, timed = de.rhs_plugins.trace_body(de.benchmarks.brusselator, 2, ("a", "b"))
    assert body == ("dy[0] = ((p[0] + ((y[0] * y[0]) * y[1])) - ((p[1] + 1.0) * y[0]));\n"
                    "dy[1] = ((p[1] * y[0]) - ((y[0] * y[0]) * y[1]));") and not timed
    assert de.rhs_plugins.trace_body(de.benchmarks.driven_pendulum, 2, ("q", "f", "w"))[1]            # uses t
    bruss = de.benchmarks.brusselator_plugin()
    n = 8192
    rng = np.random.default_rng(4)
    y0 = torch.as_tensor(rng.uniform(0.2, 3.0, (2, n)))
    b_par = torch.as_tensor(rng.uniform(1.5, 3.5, n)).cuda()
    kw = dict(t=(0.0, 8.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, constants=dict(a=1.0, b=b_par))

    def untagged(t, y, a, b):
        return torch.stack([a + y[0] ** 2 * y[1] - (b + 1.0) * y[0], b * y[0] - y[0] ** 2 * y[1]])
    u = de.OdeSystem(bruss, y0, **kw)
    v = de.OdeSystem(untagged, y0, **kw)
    for s in (u, v):
        s.method = "RK45"
        s.integrate()
    assert u._launches <= 2 < v._launches
    same = float(((u.accepted == v.accepted) & (u.rejected == v.rejected)).double().mean())
    err = float(((u[-1].y - v[-1].y).abs().amax(0) / v[-1].y.abs().amax(0).clamp_min(1.0)).max())
    print("traced Brusselator plug-in: counts identical %.5f, max rel difference %.3g" % (same, err))
    assert same >= 0.999 and err <= 1e-9
    with pytest.raises(NotImplementedError):
        de.rhs_plugins.trace_body(lambda t, y: torch.stack([torch.sign(y[0]), y[1]]), 2)


def test_user_plugin_streams_a_host_resident_ensemble():
    """The flow-controlled instantiation (dsb_erk_run_host) exists for a plug-in too: pinned host y0 in, host results out,
    bit for bit the device-resident run."""
    de = _de()
    rossler = de.benchmarks.rossler_plugin()
    n = 50000 + 3
    y0 = torch.as_tensor(_ensemble(n, seed=5))
    kw = dict(t=(0.0, 4.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a = de.OdeSystem(rossler, y0.cuda(), **kw)
    a.method = "RK45"
    a.integrate()
    h = de.OdeSystem(rossler, y0.pin_memory(), host_io=True, device="cuda:0", **kw)
    h.method = "RK45"
    h.integrate()
    assert not h[-1].y.is_cuda and torch.equal(h[-1].y, a[-1].y.cpu()) and torch.equal(h.accepted.cpu(), a.accepted.cpu())
