"""Argument-type robustness of the drop-in API on the GPU: the reference takes whatever its backend can turn into an
array (lists, numpy arrays and scalars, torch tensors on any device, 0-d tensors for times); a torch user of this engine
passes the same things.  Every spelling must give the result of the canonical call bit for bit."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _canonical(dense=False, **kw):
    de = _de()
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.tensor([1.0, 1.0, 20.0], dtype=torch.float64), t=(0.0, 1.0), dt=1e-2, rtol=1e-9,
                     atol=1e-9, dense_output=dense, **kw)
    a.method = "RK45"
    return a


@pytest.mark.parametrize("spell", ["list", "numpy", "cuda", "tuple_np_scalars", "tensor_times", "tensor_dt", "np_tols"])
def test_constructor_argument_spellings(spell):
    de = _de()
    ref = _canonical()
    ref.integrate()
    y0 = [1.0, 1.0, 20.0]
    kw = dict(t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9)
    if spell == "list":
        y = y0
    elif spell == "numpy":
        y = np.array(y0)
    elif spell == "cuda":
        y = torch.tensor(y0, dtype=torch.float64, device="cuda")
    else:
        y = torch.tensor(y0, dtype=torch.float64)
    if spell == "tuple_np_scalars":
        kw["t"] = (np.float64(0.0), np.float64(1.0))
        kw["dt"] = np.float64(1e-2)
    if spell == "tensor_times":
        kw["t"] = (torch.tensor(0.0, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64, device="cuda"))
    if spell == "tensor_dt":
        kw["dt"] = torch.tensor(1e-2, dtype=torch.float64, device="cuda")
    if spell == "np_tols":
        kw["rtol"], kw["atol"] = np.float64(1e-9), np.array(1e-9)
    a = de.OdeSystem(de.rhs_plugins.lorenz, y, **kw)
    a.method = "RK45"
    a.integrate()
    assert len(a) == len(ref) and torch.equal(a.y.cpu(), ref.y.cpu()) and torch.equal(a.t.cpu(), ref.t.cpu())


@pytest.mark.parametrize("spell", ["float", "np", "tensor0d_cpu", "tensor0d_cuda", "int"])
def test_integrate_target_time_spellings(spell):
    ref = _canonical()
    ref.integrate(0.5)
    ref.integrate(1)
    a = _canonical()
    t = dict(float=0.5, np=np.float64(0.5), tensor0d_cpu=torch.tensor(0.5, dtype=torch.float64),
             tensor0d_cuda=torch.tensor(0.5, dtype=torch.float64, device="cuda"), int=0.5)[spell]
    a.integrate(t)
    a.integrate(1 if spell == "int" else 1.0)
    assert len(a) == len(ref) and torch.equal(a.y, ref.y) and torch.equal(a.t, ref.t)


def test_dense_query_spellings():
    a = _canonical(dense=True)
    a.integrate()
    want = a.sol(0.37)
    for q in (np.float64(0.37), torch.tensor(0.37, dtype=torch.float64), torch.tensor(0.37, dtype=torch.float64, device="cuda")):
        assert torch.equal(a.sol(q).reshape(-1), want.reshape(-1))
    grid = [0.1, 0.37, 0.9]
    rows = a.sol(torch.tensor(grid, dtype=torch.float64, device="cuda"))
    for spelled in (grid, np.array(grid), torch.tensor(grid, dtype=torch.float64)):
        assert torch.equal(a.sol(spelled), rows)
    assert torch.equal(rows[1].reshape(-1), want.reshape(-1))
    assert torch.equal(a[torch.tensor(0.37, dtype=torch.float64, device="cuda")].y.reshape(-1), a[0.37].y.reshape(-1))


def test_solve_ivp_accepts_device_tensors():
    """Found by tools/differential_vs_reference.py: a CUDA `t_eval` used to fail in np.asarray."""
    de = _de()
    y0 = torch.tensor([1.0, 1.0, 20.0], dtype=torch.float64, device="cuda")
    te = [0.2, 0.5, 1.0]
    want = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.0), y0, method="RK45", t_eval=te, rtol=1e-9, atol=1e-9, first_step=1e-2)
    for spelled in (np.array(te), torch.tensor(te, dtype=torch.float64), torch.tensor(te, dtype=torch.float64, device="cuda")):
        got = de.solve_ivp(de.rhs_plugins.lorenz, (torch.tensor(0.0), np.float64(1.0)), y0, method="RK45", t_eval=spelled, rtol=1e-9,
                           atol=1e-9, first_step=np.float64(1e-2))
        assert torch.equal(got.y, want.y) and torch.equal(got.t, want.t)
    assert tuple(want.y.shape) == (3, 3) and float(want.t[-1]) == 1.0


def test_constants_spellings_and_float32_state_is_refused_clearly():
    de = _de()
    ref = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.tensor([2.0, 0.0], dtype=torch.float64), t=(0, 2.0), dt=1e-2, rtol=1e-8, atol=1e-8,
                       constants=dict(mu=1.5))
    ref.method = "RK45"
    ref.integrate()
    for mu in (np.float64(1.5), torch.tensor(1.5, dtype=torch.float64), torch.tensor(1.5, dtype=torch.float64, device="cuda"), np.array(1.5)):
        a = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.tensor([2.0, 0.0], dtype=torch.float64), t=(0, 2.0), dt=1e-2, rtol=1e-8,
                         atol=1e-8, constants=dict(mu=mu))
        a.method = "RK45"
        a.integrate()
        assert torch.equal(a.y, ref.y)
    # float32 input is not silently computed in another precision: it is promoted (documented) or refused with a message
    try:
        b = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.tensor([2.0, 0.0], dtype=torch.float32), t=(0, 2.0), dt=1e-2, rtol=1e-8,
                         atol=1e-8, constants=dict(mu=1.5))
        assert b.y.dtype == torch.float64
    except (TypeError, ValueError, NotImplementedError) as e:
        assert "float64" in str(e) or "dtype" in str(e)


def test_reference_attributes_exist_and_mean_the_same():
    """`counter`, `get_step_interpolant()`, `events_dict`, `staggered_mask`, `set_kick_vars` of the reference's OdeSystem
    (differential_system.py:523, 573-593, 769-792, 871-872)."""
    de = _de()
    a = _canonical(dense=True)
    assert a.counter == 0 and len(a) == 1
    a.integrate()
    assert a.counter == len(a) - 1 > 50 and float(a.t[a.counter]) == 1.0 == float(a.get_current_time())
    t_end, interp = a.get_step_interpolant()
    assert float(t_end) == 1.0
    mid = 0.5 * (float(a.t[-2]) + float(a.t[-1]))
    assert torch.equal(interp(mid).reshape(-1), a.sol(mid).reshape(-1))          # the last step's Hermite piece
    assert a.staggered_mask is None
    a.set_kick_vars(None)
    with pytest.raises(NotImplementedError):
        a.set_kick_vars(np.array([True, False, True]))

    def plane(t, y, **kw):
        return y[2] - 25.0

    def late(t, y, **kw):
        return t - 0.75
    b = _canonical(dense=True)
    b.integrate(events=[plane, late])
    d = b.events_dict
    assert set(d) == {e.event for e in b.events} and late in d
    assert float(d[late].t[0]) == pytest.approx(0.75, abs=1e-12) and tuple(d[late].y.shape) == (1, 3)
    assert sum(int(v.t.shape[0]) for v in d.values()) == len(b.events)
