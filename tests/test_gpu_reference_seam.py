"""Seam 1 of SURVEY.md section 8b on the GPU, with the UNMODIFIED reference in the driver's seat: the reference's own
`OdeSystem` (torch backend, CUDA tensors; baseline/_ref) gets this package's integrator classes through `set_method` and
runs ITS integrate loop (differential_system.py:996-1084) around them -- every step is one call of the engine's stage
kernels through the protocol `integrator(rhs, t, y, constants, timestep) -> (new_dt, (dTime, dState))`.
Skipped where baseline/_ref did not travel."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "desolver")),
                                 reason="baseline/_ref (the installed reference) is not present")]


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import differential_vs_reference as dv
    import desolver_b200 as de
    mod = dv._import_reference()
    de.integrators.register_with_reference(mod.integrators)
    return mod


def _lorenz(xp):
    def rhs(t, y, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
        return xp.stack([sigma * (y[1] - y[0]), y[0] * (rho - y[2]) - y[1], y[0] * y[1] - beta * y[2]])
    return rhs


@pytest.mark.parametrize("cls_name", ["RK45CKSolver", "RK8713MSolver", "DOPRI45", "RK4Solver", "HeunEulerSolver"])
def test_reference_odesystem_steps_through_the_engines_integrators(ref, cls_name):
    import desolver_b200 as de
    kw = dict(t=(0.0, 1.0), dt=0.01, rtol=1e-8, atol=1e-8)
    if cls_name == "HeunEulerSolver":
        kw.update(rtol=1e-5, atol=1e-5)
    y0 = np.array([1.0, 1.0, 20.0])
    # (a) the reference's loop around the ENGINE's integrator class, state on the GPU
    r = ref.OdeSystem(_lorenz(torch), y0=torch.as_tensor(y0).cuda(), **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # "foreign integrator" warning of :861-864
        r.set_method(getattr(de.integrators, cls_name))
    assert type(r.integrator).__module__.startswith("desolver_b200")
    r.integrate()
    assert r.y[-1].is_cuda
    # (b) the reference entirely on its own (numpy backend, its own integrator class)
    p = ref.OdeSystem(_lorenz(np), y0=y0.copy(), **kw)
    p.method = cls_name
    p.integrate()
    # (c) the engine entirely on its own
    a = de.OdeSystem(_lorenz(torch), torch.as_tensor(y0), strict=True, **kw)
    a.method = cls_name
    a.integrate()
    assert len(r) == len(p) == len(a)
    y_r, y_p, y_a = r.y[-1].cpu().numpy(), np.asarray(p.y[-1]), a.y[-1].cpu().numpy()
    # (RK8(7) from dt0 = 1e-2: the first error estimates are rounding noise, so the step SIZES ride on last-ulp details of
    # pow / atan -- numpy's against the device's -- while the step count and the result hold)
    y_bar, t_bar = (1e-9, 1e-5) if cls_name == "RK8713MSolver" else (1e-11, 1e-10)
    assert np.max(np.abs(y_r - y_p)) <= y_bar * np.max(np.abs(y_p)) and np.max(np.abs(y_a - y_p)) <= y_bar * np.max(np.abs(y_p))
    assert np.max(np.abs(np.asarray([float(v) for v in r.t]) - np.asarray(p.t, dtype=np.float64))) <= t_bar
