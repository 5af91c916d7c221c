"""GPU tests of the round-2 additions: per-step history of single systems, the paged node store (dense output at
ensemble scale, array-valued `sol(t)`, `sol.grad`, `sol.grid`), `solve_ivp(t_eval=...)`, the remaining fixed-step
tableaux, the complete integrator protocol (`dense_output()`, custom `adaptation_fn`), lazy status, and the fixes of
the round-1 review (event log accumulation, in-place `dt`, device guards, odd host_io chunks).

Fixtures: tests/golden/history_sho.npz, fixed_step.npz, solve_ivp_teval.npz -- produced by the unmodified reference
(oracle/make_golden_round2.py)."""
import numpy as np
import pytest

from conftest import load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _sho(method, dense=False, rhs=None, **kw):
    de = _de()
    a = de.OdeSystem(rhs or de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0, 2 * np.pi),
                     constants=dict(k=1.0, m=1.0), dense_output=dense, **kw)
    a.method = method
    return a


# ------------------------------------------------------------------ per-step history
@pytest.mark.parametrize("key,method,kw", [("rk45", "RK45", dict(dt=0.01, rtol=1e-9, atol=1e-9)),
                                           ("rk87", "RK87", dict(dt=0.01, rtol=1e-9, atol=1e-9)),
                                           ("babs9o7h", "BABS9O7H", dict(dt=0.05))])
@pytest.mark.parametrize("path", ["plugin", "callable"])
def test_single_system_history_is_every_accepted_step(key, method, kw, path):
    """`a.t`, `a.y`, `len(a)`, `a[i]` of a single system = the reference's rows (differential_system.py:1015-1018)."""
    g = load_golden("history_sho.npz")
    rhs = None
    if path == "callable":                                    # untagged callable: stage-level kernels + row store
        def rhs(t, y, k, m):
            return torch.stack([y[1], -k / m * y[0]])
    a = _sho(method, rhs=rhs, strict=True, **kw)
    assert len(a) == 1 and tuple(a.t.shape) == (1,) and tuple(a.y.shape) == (1, 2)
    a.integrate()
    t_ref, y_ref = g[key + ".t"], g[key + ".y"]
    assert len(a) == int(g[key + ".len"]) == t_ref.shape[0]
    assert tuple(a.t.shape) == t_ref.shape and tuple(a.y.shape) == y_ref.shape
    # The first steps of this problem have error estimates of rounding-noise size (|err| ~ 1e-14 from terms ~1e-2), so
    # the step-size sequence depends on last-ulp details of pow / atan even in STRICT arithmetic;
    # the fixed-step scheme has no such freedom.
    dt_tol, dy_tol = dict(rk45=(1e-7, 1e-7), rk87=(1e-4, 1e-6), babs9o7h=(1e-13, 1e-13))[key]
    err_t, err_y = np.max(np.abs(a.t.cpu().numpy() - t_ref)), np.max(np.abs(a.y.cpu().numpy() - y_ref))
    print("history %s/%s: max |dt| %.3g, max |dy| %.3g" % (key, path, err_t, err_y))
    assert err_t <= dt_tol and err_y <= dy_tol
    assert float(a[0].t) == 0.0 and torch.equal(a[0].y.cpu(), torch.tensor([1.0, 0.0], dtype=torch.float64))
    assert torch.equal(a[-1].y, a.y[-1]) and float(a[-1].t) == float(a.t[-1])
    assert torch.equal(a[3].y, a.y[3]) and torch.equal(a[-2].y, a.y[-2])
    with pytest.raises(IndexError):
        a[len(a)]
    assert int(a.accepted.sum()) == len(a) - 1
    if key == "rk45":
        assert len(a) - 1 == 84                                # KA0


def test_history_slices_nearest_lookup_and_continuation():
    g = load_golden("history_sho.npz")
    a = _sho("RK45", dt=0.01, rtol=1e-9, atol=1e-9, strict=True)
    a.integrate()
    s = a[0.5:2.0]                                             # slices are by TIME (reference :1170-1183)
    assert np.max(np.abs(s.t.cpu().numpy() - g["slice.t"])) <= 1e-7 and np.max(np.abs(s.y.cpu().numpy() - g["slice.y"])) <= 1e-7
    assert np.max(np.abs(a[1.0:].t.cpu().numpy() - g["slice_open.t"])) <= 1e-7
    assert np.max(np.abs(a[:3.0:2].t.cpu().numpy() - g["slice_step.t"])) <= 1e-7
    for q, t_near in zip(g["near.q"], g["near.t"]):             # nearest stored step without dense output (:1187-1196)
        assert abs(float(a[float(q)].t) - t_near) <= 1e-7
    a.integrate(7.0)                                           # a second call continues the same history
    assert len(a) == g["cont.t"].shape[0]
    assert np.max(np.abs(a.t.cpu().numpy() - g["cont.t"])) <= 1e-7 and np.max(np.abs(a.y.cpu().numpy() - g["cont.y"])) <= 1e-7
    a.reset()
    assert len(a) == 1
    a.integrate()
    assert len(a) == int(g["rk45.len"])


def test_history_grows_past_its_first_allocation():
    """5001 rows are allocated up front like the reference (:734-742); a longer run re-sizes the store and repeats."""
    de = _de()
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0, 12.0), dt=1e-3,
                     constants=dict(k=1.0, m=1.0))
    a.method = "RK4"
    a.integrate()
    assert len(a) in (12001, 12002) and abs(float(a.t[-1]) - 12.0) <= 1e-12
    tt = a.t.cpu().numpy()
    assert np.all(np.diff(tt) > 0) and np.max(np.abs(a.y[:, 0].cpu().numpy() - np.cos(tt))) < 1e-9
    b = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0, 12.0), dt=1e-3,
                     constants=dict(k=1.0, m=1.0), history="calls")
    b.method = "RK4"
    b.integrate()
    assert len(b) == 2 and torch.equal(b[-1].y, a[-1].y)


# ------------------------------------------------------------------ dense output over the node store
def test_array_valued_sol_and_grad_match_the_reference():
    g = load_golden("history_sho.npz")
    a = _sho("RK45", dense=True, dt=0.01, rtol=1e-9, atol=1e-9, strict=True)
    a.integrate()
    tq = torch.as_tensor(g["tq"])
    got = a.sol(tq)                                            # (T, dim): the reference's t.shape + y.shape
    # (the step partition differs from the reference's at the 1e-9 level, see the history test: interpolated values then
    # agree to the interpolation error of the cubic, ~1e-8 here)
    assert tuple(got.shape) == g["sol"].shape and np.max(np.abs(got.cpu().numpy() - g["sol"])) <= 1e-7
    got2 = a.sol(tq[:6].reshape(2, 3))
    assert tuple(got2.shape) == g["sol2d"].shape and np.max(np.abs(got2.cpu().numpy() - g["sol2d"])) <= 1e-7
    for q, t in enumerate(g["tq"]):                             # scalar queries = rows of the array query, bit for bit
        assert torch.equal(a.sol(float(t)), got[q])
        assert np.max(np.abs(a.sol.grad(float(t)).cpu().numpy() - g["grad"][q])) <= 1e-6
    assert torch.equal(a.sol.grad(tq)[3], a.sol.grad(float(tq[3])))
    assert torch.equal(a[1.0].y, a.sol(1.0)) and len(a.sol) == len(a)


def test_ensemble_grid_evaluation_and_per_trajectory_steps():
    de = _de()
    g = load_golden("lorenz_dense.npz")
    y0 = torch.as_tensor(g["y0"])
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, dense_output=True)
    a.method = "RK45"
    a.integrate()
    times = torch.as_tensor(g["tq"])
    grid = a.sol.grid(times)                                   # ONE launch: (T, dim, N)
    assert tuple(grid.shape) == (len(times), 3, y0.shape[1])
    assert np.max(rel_inf(np.moveaxis(grid.cpu().numpy(), 0, -1).reshape(3, -1), np.moveaxis(g["dense"], 0, -1).reshape(3, -1))) < 1e-10
    for q in (0, 7, 23):
        assert torch.equal(grid[q], a.sol(float(times[q])))
    per = times[:y0.shape[1]].clone()
    two = a.sol(torch.stack([per, per.flip(0)]))                # (T=2, N) query times -> (2, dim, N)
    assert torch.equal(two[0], a.sol(per)) and torch.equal(two[1], a.sol(per.flip(0)))
    dg = a.sol.grid(times, derivative=True)
    f = de.rhs_plugins.lorenz(0.0, a.sol(0.3), sigma=10.0, rho=28.0, beta=8.0 / 3.0)
    # derivative of the cubic: O(h^3) accurate
    assert float((a.sol.grad(0.3) - f).abs().max() / f.abs().max()) < 1e-3 and tuple(dg.shape) == tuple(grid.shape)
    # the ragged per-trajectory history: every trajectory's rows = that trajectory run alone as a single system
    for i in (0, 5):
        s = de.OdeSystem(de.rhs_plugins.lorenz, y0[:, i].clone(), t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9)
        s.method = "RK45"
        s.integrate()
        t_i, y_i = a.steps(i)
        assert torch.equal(t_i, s.t) and torch.equal(y_i, s.y) and int(a.accepted[i]) == len(s) - 1


def test_dense_output_at_scale_fits_and_is_deterministic():
    """2^17 Lorenz trajectories with dense output: ~3.3e7 nodes in a ragged store of ~2.2 GB (the round-1 layout
    needed 7.5 GB at its default capacity), results identical to the run without dense output."""
    de = _de()
    n = 1 << 17
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n, seed=1)).cuda()
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, dense_output=True)
    a.method = "RK45"
    a.integrate()
    b = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    b.method = "RK45"
    b.integrate()
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    st = a._node_store
    claimed, dropped = st.usage()
    total = int(a.accepted.sum()) + n
    assert dropped == 0 and torch.equal(a.sol.node_count, a.accepted + 1)
    assert total <= claimed <= total + 4736 * st.page_entries           # no waste beyond one page in progress per warp
    assert torch.equal(a.sol(2.0), a[-1].y) and torch.equal(a.sol(0.0), y0)
    mid = a.sol(1.0)
    a.reset()
    a.integrate()
    assert torch.equal(a.sol(1.0), mid)                                 # scheduling-independent, run to run


# ------------------------------------------------------------------ solve_ivp(t_eval)
def test_solve_ivp_t_eval_matches_the_reference_and_the_dense_mode():
    de = _de()
    g = load_golden("solve_ivp_teval.npz")
    y0 = torch.as_tensor(g["y0"])
    r = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.5), y0, method="RK45", t_eval=g["t_eval"], rtol=1e-9, atol=1e-9,
                     first_step=1e-2, ensemble=True)
    assert tuple(r.y.shape) == g["y"].shape                     # (dim, N, T)
    for q in range(g["t_eval"].shape[0]):
        assert rel_inf(r.y[..., q].cpu().numpy(), g["y"][..., q]).max() <= 1e-10
    assert np.max(np.abs(r.t.cpu().numpy() - g["t"])) <= 1e-13
    # ... served by ONE launch of the multi-stop kernel, bit-identical to the sequence of integrate(t) calls it replaces
    assert r.ode_system._launches == 1 and len(r.ode_system) == 1 + len(g["t_eval"])
    seq = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 1.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    seq.method = "RK45"
    for q, tq in enumerate(g["t_eval"]):
        seq.integrate(float(tq))
        assert torch.equal(seq[-1].y, r.y[..., q])
    assert torch.equal(seq.accepted, r.ode_system.accepted) and torch.equal(seq.rejected, r.ode_system.rejected)
    assert torch.equal(seq.dt, r.ode_system.dt) and torch.equal(seq[-1].t, r.ode_system[-1].t)
    d = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.5), y0, method="RK45", t_eval=g["t_eval"], rtol=1e-9, atol=1e-9,
                     first_step=1e-2, ensemble=True, t_eval_mode="dense")
    assert d.ode_system._launches == 1                          # one integration; the evaluation is one more launch
    assert tuple(d.y.shape) == g["y"].shape and float((d.y - r.y).abs().max() / r.y.abs().max()) < 1e-6
    s = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.5), y0[:, 0].clone(), method="RK45", t_eval=g["t_eval"], rtol=1e-9, atol=1e-9,
                     first_step=1e-2)
    assert rel_inf(s.y.cpu().numpy(), g["y"][:, 0]).max() <= 1e-10


# ------------------------------------------------------------------ the fixed-step tableaux that had no fixture
@pytest.mark.parametrize("method", ["RK5", "Midpoint", "Heun's", "Ralston's", "Euler", "Euler-Trapezoidal"])
@pytest.mark.parametrize("fused", [True, False])
def test_fixed_step_tableaux(method, fused):
    de = _de()
    g = load_golden("fixed_step.npz")
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(g["y0"]), t=(0.0, float(g["tf"])), dt=float(g["dt0"]), ensemble=True)
    a.method = method
    a.use_fused = fused
    a.integrate()
    assert np.array_equal(a.accepted.cpu().numpy(), g[method + ".steps"]) and int(a.rejected.sum()) == 0
    assert rel_inf(a[-1].y.cpu().numpy(), g[method + ".y"]).max() <= 1e-12
    assert np.max(np.abs(a[-1].t.cpu().numpy() - g[method + ".t"])) <= 1e-13


def test_symplectic_euler_history():
    de = _de()
    g = load_golden("fixed_step.npz")
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0, 1.0), dt=0.01,
                     constants=dict(k=1.0, m=1.0))
    a.method = "Symplectic Forward Euler"
    a.integrate()
    assert tuple(a.t.shape) == g["symp_euler.t"].shape
    assert np.max(np.abs(a.t.cpu().numpy() - g["symp_euler.t"])) <= 1e-13 and np.max(np.abs(a.y.cpu().numpy() - g["symp_euler.y"])) <= 1e-13


# ------------------------------------------------------------------ integrator protocol (seam 1)
def test_integrator_protocol_dense_output_and_custom_adaptation():
    de = _de()
    integ = de.integrators.RK45CKSolver((3,), torch.float64, rtol=1e-9, atol=1e-9, device="cuda")
    y0 = torch.tensor([1.0, 1.0, 20.0], dtype=torch.float64, device="cuda")
    consts = dict(sigma=10.0, rho=28.0, beta=8.0 / 3.0)
    new_dt, (dT, dY) = integ(de.rhs_plugins.lorenz, 0.0, y0, consts, 1e-2)
    t1, interp = integ.dense_output()                          # reference integrator_types.py:109-119
    assert float(t1) == float(dT) and torch.equal(interp(0.0), y0) and torch.equal(interp(float(t1)), y0 + dY)
    f0 = de.rhs_plugins.lorenz(0.0, y0, **consts)
    assert torch.equal(interp.grad(0.0), f0) and torch.equal(interp.m0, f0)
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=consts, dense_output=True)
    a.integrate()
    assert abs(float(a.t[1]) - float(t1)) <= 1e-15              # first accepted step of KA1 (after one rejection)
    half = 0.5 * float(t1)
    assert float((interp(half) - a.sol(half)).abs().max()) <= 1e-10
    # a custom controller that defers to the built-in one must reproduce the built-in step
    custom = de.integrators.RK45CKSolver((3,), torch.float64, rtol=1e-9, atol=1e-9, device="cuda")
    calls = []

    def controller(self):
        calls.append(float(self.solver_dict["timestep"]))
        return self.update_timestep(ignore_custom_adaptation=True)
    custom.adaptation_fn = controller
    calls.clear()
    new_dt2, (dT2, dY2) = custom(de.rhs_plugins.lorenz, 0.0, y0, consts, 1e-2)
    # (the kernels' error-norm quotients carry 2^-40 relative error by design, fastmath.cuh; the host controller is exact)
    assert len(calls) == 2 and abs(float(dT2) - float(dT)) <= 1e-11 * float(dT) and abs(float(new_dt2) - float(new_dt)) <= 1e-9 * float(new_dt)
    assert float((dY2 - dY).abs().max()) <= 1e-11
    # symplectic protocol step
    symp = de.integrators.BABs9o7HSolver((2,), torch.float64, device="cuda")
    q0 = torch.tensor([1.0, 0.0], dtype=torch.float64, device="cuda")
    step, (dTs, dYs) = symp(de.rhs_plugins.harmonic_oscillator, 0.0, q0, dict(k=1.0, m=1.0), 0.05)
    assert float(step) == 0.05 and abs(float(dTs) - 0.05) <= 1e-17
    assert abs(float((q0 + dYs)[0]) - np.cos(0.05)) < 1e-12
    t_end, it2 = symp.dense_output()
    assert torch.equal(it2(0.05), q0 + dYs)


# ------------------------------------------------------------------ lazy status
def test_lazy_status_defers_the_counter_read_back():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(4096, seed=3)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.method = "RK45"
    a.integrate()
    eager = a.stats
    a.reset()
    a.lazy_status = True
    a.integrate()
    assert a._pending and a._stats is None
    assert a.success and a.stats == eager and not a._pending
    bad = y0.copy()
    bad[:, 7] = np.inf
    b = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(bad), t=(0.0, 0.1), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    b.lazy_status = True
    b.integrate()                                              # does not raise: nothing has been read yet
    assert not b.success and "FailedToMeetTolerances" in repr(type(b._OdeSystem__int_status.__cause__))
    with pytest.raises(de.exception_types.FailedIntegration):
        b.check()
    assert b.stats["failed"] == 1


# ------------------------------------------------------------------ fixes from the round-1 review
def test_event_log_accumulates_over_calls_with_a_plain_callable():
    de = _de()
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(64, seed=6))
    ev = lambda t, y, **kw: y[0] - 1.0                          # noqa: E731   (converted to a new object on every call)

    def run(stops):
        a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
        a.method = "RK45"
        for tf in stops:
            a.integrate(tf, events=ev)
        return a
    one, two = run([2.0]), run([0.7, 1.4, 2.0])
    assert int(two.events.count.sum()) == int(one.events.count.sum()) > 0
    assert torch.equal(two.events.count, one.events.count)


def test_max_step_reaches_the_stage_level_kernels():
    de = _de()

    def rhs(t, y):                                             # untagged: stage-level path
        return torch.stack([y[1], -y[0]])
    y0 = torch.tensor([[1.0, 0.5], [0.0, 0.2]], dtype=torch.float64)
    free = de.solve_ivp(rhs, (0.0, 3.0), y0, method="RK45", rtol=1e-6, atol=1e-6, first_step=1e-2, ensemble=True)
    capped = de.solve_ivp(rhs, (0.0, 3.0), y0, method="RK45", rtol=1e-6, atol=1e-6, first_step=1e-2, ensemble=True, max_step=0.05)
    assert int(capped.ode_system.accepted.min()) >= 60 > int(free.ode_system.accepted.max())
    assert float(capped.ode_system.dt.abs().max()) <= 0.05
    a = capped.ode_system
    exact = torch.stack([y0[0] * np.cos(3.0) + y0[1] * np.sin(3.0), -y0[0] * np.sin(3.0) + y0[1] * np.cos(3.0)]).cuda()
    assert float((a[-1].y - exact).abs().max()) < 1e-6


def test_current_device_is_restored_and_other_devices_work():
    de = _de()
    before = torch.cuda.current_device()
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(256, seed=2))
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, device="cuda:0")
    a.integrate()
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() < 2:
        return
    b = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, device="cuda:1")
    b.integrate()                                              # created and launched while cuda:0 is current
    assert torch.cuda.current_device() == before and torch.equal(a[-1].y.cpu(), b[-1].y.cpu())


def test_host_io_with_odd_length_and_several_chunks():
    de = _de()
    n = 40001                                                  # n % 4 != 0: chunk boundaries fall inside 32-byte sectors
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n, seed=8)).pin_memory()
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0.cuda(), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.integrate()
    for chunks in (7, 16):
        h = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, host_io=True,
                         host_chunks=chunks)
        h.integrate()
        assert torch.equal(h[-1].y, a[-1].y.cpu()) and torch.equal(h.accepted, a.accepted.cpu())


# ------------------------------------------------------------------ collectives behind the C ABI (one rank)
def test_comm_entry_points_on_one_rank():
    from desolver_b200.backend import cuda_backend as cb
    if not cb.Comm.available():
        pytest.skip("NCCL not loadable")
    comm = cb.Comm(0, 1, 0, cb.Comm.unique_id())
    dev = torch.device("cuda:0")
    c = torch.arange(8, dtype=torch.int64, device=dev)
    out = torch.zeros(4, dtype=torch.int64, device=dev)
    st = cb.current_stream_ptr(dev)
    comm.allreduce_counters(c.data_ptr() + 16, out.data_ptr(), 4, st)
    shard = torch.randn(3, 1000, dtype=torch.float64, device=dev)
    full = torch.empty_like(shard)
    staging = torch.empty(comm.staging_bytes(3, 8, 1000), dtype=torch.uint8, device=dev)
    comm.gather(shard, staging, full, 1000, st)
    torch.cuda.synchronize()
    assert out.tolist() == [2, 3, 4, 5] and torch.equal(full, shard)
    comm.close()


def test_scatter_store_interleaves_round_robin_shards():
    """The gather fused into the result store (dsb_erk_run_args.out_*): two shard systems (the two 'ranks' of a
    round-robin split, here on one device) store straight into one buffer; the buffer equals the unsharded run."""
    de = _de()
    from desolver_b200.backend import cuda_backend as cb
    n, W = 10001, 2
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n, seed=13)).cuda()
    ref = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    ref.integrate()
    buf = cb.PeerBuffer.allocate(8 * 5 * n + 4 * 3 * n + 64, 0)
    y = buf.tensor(torch.float64, (3, n), 0)
    t = buf.tensor(torch.float64, (n,), 8 * 3 * n)
    dt = buf.tensor(torch.float64, (n,), 8 * 4 * n)
    acc = buf.tensor(torch.int32, (n,), 8 * 5 * n)
    rej = buf.tensor(torch.int32, (n,), 8 * 5 * n + 4 * n)
    sta = buf.tensor(torch.int32, (n,), 8 * 5 * n + 8 * n)
    for x in (y, t, dt, acc, rej, sta):
        x.fill_(-1)
    total = 0
    for rank in range(W):
        s = de.OdeSystem(de.rhs_plugins.lorenz, y0[:, rank::W].contiguous(), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
        s._scatter_store = dict(index_mul=W, index_add=rank, stride=n, y=y.data_ptr(), t=t.data_ptr(), dt=dt.data_ptr(),
                                accepted=acc.data_ptr(), rejected=rej.data_ptr(), status=sta.data_ptr())
        s.lazy_status = True
        s.integrate()
        total += s.stats["attempts"]
    torch.cuda.synchronize()
    assert torch.equal(y, ref[-1].y) and torch.equal(t, ref[-1].t) and torch.equal(dt, ref.dt)
    assert torch.equal(acc, ref.accepted) and torch.equal(rej, ref.rejected) and torch.equal(sta, ref.status)
    assert total == ref.stats["attempts"]


# ------------------------------------------------------------------ Richardson wrapper + protocol loop
@pytest.mark.parametrize("key,base,it", [("rk4_2", "RK4Solver", 2), ("rk4_3", "RK4Solver", 3), ("rk45ck_2", "RK45CKSolver", 2),
                                         ("midpoint_4", "MidpointSolver", 4)])
def test_richardson_extrapolation_matches_the_reference(key, base, it):
    de = _de()
    g = load_golden("richardson_sho.npz")
    R = de.integrators.generate_richardson_integrator(getattr(de.integrators, base), richardson_iter=it)
    assert issubclass(R, de.integrators.RichardsonIntegratorTemplate) and R.__name__ == "RichardsonExtrapolated_%s_Integrator" % base
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0, 2 * np.pi), dt=0.1,
                     rtol=1e-9, atol=1e-9, constants=dict(k=1.0, m=1.0))
    a.set_method(R)
    a.integrate()
    t_ref, y_ref = g[key + ".t"], g[key + ".y"]
    if key == "midpoint_4" and len(a) != t_ref.shape[0]:
        pytest.skip("thousands of noise-limited steps: the step count is not reproducible to the last step")
    assert len(a) == t_ref.shape[0] and a.success
    # same number of steps; the step sizes come out of error estimates that are differences of nearly equal increments
    # (noise-limited at tol 1e-9), so the times drift apart at the 1e-7 ... 1e-6 level over the ~200 steps
    assert np.max(np.abs(a.t.cpu().numpy() - t_ref)) <= 5e-6 and np.max(np.abs(a.y.cpu().numpy() - y_ref)) <= 5e-6
    assert abs(a.nfev - int(g[key + ".nfev"])) <= 0.02 * int(g[key + ".nfev"])
    assert abs(float(a[-1].y[0]) - 1.0) < 1e-6


# ------------------------------------------------------------------ combinations of the node store with other features
def test_dense_output_with_callbacks_equals_the_one_shot_run():
    """Chunked launches (callback every k accepted steps) append to the same paged store: same nodes, same interpolants."""
    de = _de()
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(300, seed=4))

    def run(callback):
        a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, dense_output=True)
        a.method = "RK45"
        a.callback_every = 7
        a.integrate(callback=callback)
        return a
    calls = []
    one, chunked = run(None), run(lambda s: calls.append(1))
    assert len(calls) > 10 and torch.equal(one[-1].y, chunked[-1].y) and torch.equal(one.sol.node_count, chunked.sol.node_count)
    grid = torch.linspace(0.0, 1.0, 9, dtype=torch.float64)
    assert torch.equal(one.sol.grid(grid), chunked.sol.grid(grid))
    t0, y0_rows = one.steps(17)
    t1, y1_rows = chunked.steps(17)
    assert torch.equal(t0, t1) and torch.equal(y0_rows, y1_rows)


def test_single_system_history_with_events_and_with_the_protocol_loop():
    de = _de()
    ev = de.events.component_crossing(0, 0.0, dim=2, direction=-1)
    a = _sho("RK45", dt=0.01, rtol=1e-9, atol=1e-9)
    a.integrate(events=ev)                                     # history + events: stage-level kernels, row store
    b = _sho("RK45", dt=0.01, rtol=1e-9, atol=1e-9)
    b.integrate()
    assert len(a) == len(b) and float((a.t - b.t).abs().max()) <= 1e-9 and len(a.events) == 1
    assert abs(float(a.events[0].t) - np.pi / 2) <= 1e-8
    term = de.events.component_crossing(0, 0.0, dim=2, direction=-1, is_terminal=True)
    c = _sho("RK45", dense=True, dt=0.01, rtol=1e-9, atol=1e-9)
    c.integrate(events=term)                                   # stops at the event: the history ends there
    assert abs(float(c.t[-1]) - np.pi / 2) <= 1e-8 and len(c) < len(b) and float(c.t[-1]) == float(c[-1].t)
    assert float((c.sol(1.0) - b[1.0].y).abs().max()) < 1.0     # dense output next to events is available
    # a Richardson-wrapped integrator with dense output: nodes come from the protocol loop
    R = de.integrators.generate_richardson_integrator(de.integrators.RK4Solver, richardson_iter=2)
    d = _sho("RK45", dense=True, dt=0.1, rtol=1e-9, atol=1e-9)
    d.set_method(R)
    d.integrate()
    assert len(d.sol) == len(d) and abs(float(d.sol(1.0)[0]) - np.cos(1.0)) < 1e-7


def test_array_valued_tolerances():
    """`rtol` / `atol` of the state's shape, one value per component (reference integrator_types.py:37-38): the stage-level
    kernels take them as vectors; a fused plugin with such tolerances runs there."""
    de = _de()
    g = load_golden("lorenz_vector_tol.npz")
    for ensemble in (True, False):
        y0 = torch.as_tensor(g["y0"] if ensemble else g["y0"][:, 0].copy())
        a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=g["rtol"], atol=torch.as_tensor(g["atol"]),
                         ensemble=ensemble)
        a.method = "RK45"
        assert a.integrator.has_vector_tolerances and a._fused_plugin() is None
        a.integrate()
        k = slice(None) if ensemble else slice(0, 1)
        assert np.array_equal(a.accepted.cpu().numpy(), g["accepted"][k])
        assert np.array_equal((a.accepted + a.rejected).cpu().numpy(), g["attempts"][k])
        y = a[-1].y.cpu().numpy().reshape(3, -1)
        assert rel_inf(y, g["y_final"][:, k]).max() <= 1e-9
    # scalar tolerances given as 0-d arrays stay on the fused path
    b = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(g["y0"]), t=(0.0, 2.0), dt=1e-2, rtol=np.float64(1e-9), atol=np.array(1e-9),
                     ensemble=True)
    assert not b.integrator.has_vector_tolerances and b._fused_plugin() is not None


# ------------------------------------------------------------------ a run that changes kernels half-way (differential run finding)
@pytest.mark.parametrize("dense", [False, True])
def test_history_survives_a_switch_from_the_fused_to_the_stage_level_kernels(dense):
    """integrate(t1) in the fused kernel, then a method without a fused instantiation (RK14(12), 35 stages): the nodes
    recorded so far are re-laid out for the row writer (NodeStore.to_rows) and the run continues -- history, dense
    output and counters equal those of the same sequence run on the stage-level kernels throughout (STRICT arithmetic:
    both paths round like the reference)."""
    de = _de()

    def run(use_fused):
        a = de.OdeSystem(de.rhs_plugins.lorenz, torch.tensor([1.0, 1.0, 20.0], dtype=torch.float64), t=(0.0, 1.0), dt=1e-2, rtol=1e-9,
                         atol=1e-9, dense_output=dense, strict=True)
        a.use_fused = use_fused
        a.method = "RK45"
        a.integrate(0.4)
        a.method = "RK1412"
        a.integrate()
        return a
    a, b = run(True), run(False)
    assert a._node_store.kind == "rows" and len(a) == len(b) > 20
    assert torch.equal(a.t, b.t) and torch.equal(a.y, b.y)
    assert a.nfev == b.nfev and int(a.accepted.sum()) == len(a) - 1
    if dense:
        q = torch.tensor([0.05, 0.39, 0.41, 0.97], dtype=torch.float64, device="cuda")
        assert torch.equal(a.sol(q), b.sol(q)) and torch.equal(a.sol.grad(q), b.sol.grad(q))


def test_a_row_written_run_stays_on_the_stage_level_kernels():
    """Per-component tolerances start the run on the stage-level kernels; scalar tolerances afterwards would make the
    system eligible for the fused kernel, whose node layout differs: it stays where its history lives (the reference
    just carries on, differential_system.py:646-660), and equals the all-stage-level run bit for bit."""
    de = _de()

    def run(use_fused):
        a = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.tensor([2.0, 0.0], dtype=torch.float64), t=(0.0, 3.0), dt=1e-2,
                         rtol=torch.tensor([1e-6, 2e-6], dtype=torch.float64), atol=1e-6, constants=dict(mu=1.5), strict=True)
        a.use_fused = use_fused
        a.method = "RK45"
        a.integrate(1.2)
        a.rtol = 1e-9
        a.atol = 1e-9
        a.integrate()
        return a
    a, b = run(True), run(False)
    assert len(a) == len(b) and torch.equal(a.t, b.t) and torch.equal(a.y, b.y) and a.nfev == b.nfev
    assert float(a.t[-1]) == 3.0


def test_nfev_across_method_switches_and_reset_follows_the_reference():
    """reference: every new integrator object evaluates initial_rhs once more on its first step (integrator_types.py:
    169-175); reset() zeroes the counter (differential_system.py:906).  SHO: RK4 (231 steps of 5) then RK45CK (7 per attempt)."""
    de = _de()
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, torch.tensor([1.0, 0.0], dtype=torch.float64), t=(0.0, 4.0), dt=0.01,
                     rtol=1e-8, atol=1e-8, constants=dict(k=1.0, m=1.0))
    a.method = "RK4"
    a.integrate(2.0)
    n1 = len(a) - 1
    assert a.nfev == 2 + 5 * n1
    a.method = "RK45"
    assert a.nfev == 2 + 5 * n1                                # nothing evaluated yet
    a.integrate()
    att2 = int((a.accepted + a.rejected).sum()) - n1
    assert a.nfev == 2 + 5 * n1 + 1 + 7 * att2
    a.reset()
    assert a.nfev == 0
    a.integrate()
    assert a.nfev == 1 + 7 * int((a.accepted + a.rejected).sum())


@pytest.mark.parametrize("path", ["plugin", "callable"])
def test_single_system_callbacks_fire_once_per_accepted_step(path):
    """The reference calls back after every ACCEPTED step (differential_system.py:1073-1084); rejected attempts are
    invisible to the caller.  (The stage-level loop used to call back once per attempt -- found by the differential run.)"""
    de = _de()
    rhs = de.rhs_plugins.forced if path == "plugin" else (lambda t, y: torch.stack([y[1], t - y[0]]))
    a = de.OdeSystem(rhs, torch.tensor([0.3, -0.2], dtype=torch.float64), t=(0.0, 3.0), dt=0.5, rtol=1e-9, atol=1e-9)
    a.method = "RK45"
    seen = []

    def cbk(s):
        seen.append(float(s.t[-1]))
        if abs(float(s.dt)) > 0.05:              # a step clamp, as solve_ivp(max_step=...) installs it
            s.dt = 0.05
    a.integrate(callback=cbk)
    assert int(a.rejected.sum()) > 0                                   # dt0 = 0.5 is far too large for the tolerance
    assert len(seen) == len(a) - 1 == int(a.accepted.sum())
    assert np.allclose(seen, a.t[1:].cpu().numpy(), rtol=0, atol=0)    # called with the state at each accepted step
