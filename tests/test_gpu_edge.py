"""Edge cases of the ensemble engine on the GPU: ragged sizes, empty work, shifted/negative time spans,
backward time, per-trajectory constants of a fused plugin, method switching, reset, dense-store overflow."""
import os

import numpy as np
import pytest

from conftest import load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _oracle(y0, t0, tf, dt0, tol, cls="RK45CKSolver", params=(10.0, 28.0, 8.0 / 3.0)):
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    return orc.erk_ensemble(orc.RHS_LORENZ, y0, t0, tf, dt0, tol, tol, *tb.rk_arrays(cls), params=np.asarray(params),
                            threads=os.cpu_count())


@pytest.mark.parametrize("n", [1, 31, 33, 777])
def test_ragged_ensemble_sizes(n):
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(n, seed=n)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 0.7), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.integrate()
    ref = _oracle(y0, 0.0, 0.7, 1e-2, 1e-9)
    assert np.array_equal(a.accepted.cpu().numpy(), ref["accepted"]) and np.array_equal(a.rejected.cpu().numpy(), ref["rejected"])
    assert rel_inf(a[-1].y.cpu().numpy(), ref["y"]).max() <= 1e-10


def test_nothing_to_do_and_shifted_spans():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(64, seed=1)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(1.5, 1.5 + 1e-16), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.integrate()                                   # |tf - t0| < 4 eps: integrate() returns at once (reference :956-957)
    assert int(a.accepted.sum()) == 0 and torch.equal(a[-1].y.cpu(), torch.as_tensor(y0)) and a.success
    # (forward spans that END at a smaller |t| than they start, e.g. (-1.0, -0.2), make the reference's final-step
    #  test |dt + t| > |tf| (:997) true on every step -- a reference pathology the engine mirrors; not exercised here)
    for t0, tf in ((3.0, 3.9), (-0.4, 0.5), (0.25, 1.0)):
        b = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(t0, tf), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
        b.integrate()
        ref = _oracle(y0, t0, tf, 1e-2, 1e-9)
        assert bool((b[-1].t == tf).all()) == bool((ref["t"] == tf).all())
        assert np.array_equal(b.accepted.cpu().numpy(), ref["accepted"])
        assert rel_inf(b[-1].y.cpu().numpy(), ref["y"]).max() <= 1e-10
        assert np.array_equal(b[-1].t.cpu().numpy(), ref["t"])


def test_backward_time_with_rejections():
    """tf < t0 with an oversized first step.  The reference retries such a step with the SAME h 64 times and
    raises; engine and oracle use sign(h0)*min(|new|,|h0|), so the run completes.  Negative
    times are used so that the final-step test |dt + t| > |tf| (:997) does not fire at once, and a harmonic
    oscillator because Lorenz integrated backwards amplifies last-ulp differences into different step counts."""
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    de = _de()
    rng = np.random.default_rng(6)
    y0 = rng.uniform(-2, 2, size=(2, 48))
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.as_tensor(y0), t=(-0.1, -6.1), dt=3.0, rtol=1e-9, atol=1e-9,
                     constants=dict(k=1.0, m=1.0), ensemble=True)
    a.integrate()
    ref = orc.erk_ensemble(orc.RHS_SHO, y0, -0.1, -6.1, 3.0, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"), params=[1.0, 1.0])
    assert int(a.rejected.sum()) > 0 and bool((a.dt < 0).all())
    assert np.array_equal(a.accepted.cpu().numpy(), ref["accepted"]) and np.array_equal(a.rejected.cpu().numpy(), ref["rejected"])
    assert rel_inf(a[-1].y.cpu().numpy(), ref["y"]).max() <= 1e-10 and bool((a[-1].t == -6.1).all())
    exact = np.stack([y0[0] * np.cos(-6.0) + y0[1] * np.sin(-6.0), -y0[0] * np.sin(-6.0) + y0[1] * np.cos(-6.0)])
    assert np.max(np.abs(a[-1].y.cpu().numpy() - exact)) < 1e-6


def test_per_trajectory_constants_of_a_fused_plugin():
    de = _de()
    n = 96
    y0 = de.benchmarks.lorenz_ensemble(n, seed=8)
    rho = np.linspace(20.0, 35.0, n)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9,
                     constants=dict(sigma=10.0, rho=torch.as_tensor(rho), beta=8.0 / 3.0), ensemble=True)
    a.integrate()
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)])
    ref = orc.erk_ensemble(orc.RHS_LORENZ, y0, 0.0, 1.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"), params=prm)
    assert np.array_equal(a.accepted.cpu().numpy(), ref["accepted"])
    assert rel_inf(a[-1].y.cpu().numpy(), ref["y"]).max() <= 1e-10
    with pytest.raises((ValueError, RuntimeError)):     # wrong length: caught by the constructor's RHS shape probe
        de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), constants=dict(rho=torch.ones(n + 1)), ensemble=True).integrate()


def test_method_switch_reset_and_properties():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(40, seed=12)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    assert a.integration_status == "Integration has not been run." and not a.success and len(a) == 1
    a.integrate(0.5)
    a.method = "RK87"                                # continue with another scheme from (t, y, dt)
    assert a.method is de.integrators.RK8713MSolver and a.integrator.order == 8.0
    a.integrate()
    r1 = _oracle(y0, 0.0, 0.5, 1e-2, 1e-9)
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    r2 = orc.erk_ensemble(orc.RHS_LORENZ, r1["y"], r1["t"], 1.0, r1["dt"], 1e-9, 1e-9, *tb.rk_arrays("RK8713MSolver"),
                          params=[10.0, 28.0, 8.0 / 3.0])
    assert rel_inf(a[-1].y.cpu().numpy(), r2["y"]).max() <= 1e-10 and len(a) == 3
    assert tuple(a.y.shape) == (3, 3, 40) and tuple(a.t.shape) == (3, 40) and tuple(a.dt.shape) == (40,)
    a.reset()
    assert len(a) == 1 and int(a.accepted.sum()) == 0 and torch.equal(a[0].y.cpu(), torch.as_tensor(y0))
    a.rtol = 1e-6
    a.atol = 1e-6
    a.integrate()
    assert a.integrator.rtol == 1e-6 and a.success
    assert int(a.accepted.max()) < int(r1["accepted"].max() + r2["accepted"].max())
    a.tf = 1.25
    a.integrate()
    assert bool((a[-1].t == 1.25).all())
    with pytest.raises(ValueError):
        a.tf = a.t0


def test_dense_store_overflow_and_continuation():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(16, seed=2)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True,
                     dense_output=True, dense_capacity=256)
    a.integrate(0.4)
    n1 = a.sol.node_count.clone()
    a.integrate()                                   # second call appends to the same node store
    assert bool((a.sol.node_count > n1).all())
    assert torch.equal(a.sol.node_count - 1, a.accepted)
    b = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True,
                     dense_output=True, dense_capacity=256)
    b.integrate()
    for tq in (0.1, 0.4, 0.77):
        # different step partitions (a stops at t = 0.4): agreement to the cubic-Hermite interpolation error
        assert rel_inf(a.sol(tq).cpu().numpy(), b.sol(tq).cpu().numpy()).max() < 1e-6
    c = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True,
                     dense_output=True, dense_capacity=8)
    # a pool that is too small: the kernel still counts every node, the run is repeated with a pool of the size that
    # is then known -- same results, bit for bit
    c.integrate()
    assert torch.equal(c[-1].y, b[-1].y) and torch.equal(c.accepted, b.accepted)
    assert torch.equal(c.sol.node_count, b.sol.node_count)
    assert torch.equal(c.sol(0.5), b.sol(0.5)) and torch.equal(c.sol(0.123), b.sol(0.123))
    c.integrate(1.5)                                # ... also when the overflowing call CONTINUES an integration
    b.integrate(1.5)
    assert torch.equal(c[-1].y, b[-1].y) and torch.equal(c.sol(1.2), b.sol(1.2))
    assert a.sol is not None and de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), ensemble=True).sol is None


def test_history_rows_survive_later_integrations_and_resets():
    """`a[-1]` aliases the live state until the state moves on: rows handed out earlier must keep their values when
    integrate() continues or the system is reset and run again."""
    import desolver_b200 as de
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(512, seed=9))
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.integrate(0.5)
    row1_y, row1_t = a[-1].y, a[-1].t
    keep_y, keep_t = row1_y.clone(), row1_t.clone()
    a.integrate(1.0)
    assert torch.equal(row1_y, keep_y) and torch.equal(row1_t, keep_t) and torch.equal(a[1].y, keep_y)
    assert float(a[-1].t.min()) == 1.0 and not torch.equal(a[-1].y, keep_y)
    end_y = a[-1].y
    keep_end = end_y.clone()
    a.reset()
    a.integrate(0.25)
    assert torch.equal(end_y, keep_end) and float(a[-1].t.max()) == 0.25 and len(a) == 2
