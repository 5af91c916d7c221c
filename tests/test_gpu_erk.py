"""GPU parity tests of the fused explicit-RK ensemble engine (K1): the CUDA path, called through the
public `OdeSystem` API and the C ABI underneath, against (a) the golden fixtures produced by the
unmodified reference and (b) the CPU oracle on fresh seeded inputs.

Bars (BASELINE.json north_star): final fp64 states within 1e-10 relative (norm-relative per
trajectory), accepted/rejected counts identical for >= 99.9 % of trajectories.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

STATE_TOL = 1e-10
COUNT_FRACTION = 0.999


def _de():
    import desolver_b200 as de
    return de


def _run(rhs, y0, tf, dt0, tol, method, constants=None, strict=False, t0=0.0, **kw):
    de = _de()
    a = de.OdeSystem(rhs, torch.as_tensor(np.ascontiguousarray(y0)), t=(t0, tf), dt=dt0, rtol=tol, atol=tol,
                     constants=constants or {}, ensemble=True, strict=strict, **kw)
    a.method = method
    a.integrate()
    torch.cuda.synchronize()
    return a


def _compare(a, y_ref, acc_ref, att_ref, state_tol=STATE_TOL, count_fraction=COUNT_FRACTION):
    y = a[-1].y.cpu().numpy()
    acc = a.accepted.cpu().numpy()
    rej = a.rejected.cpu().numpy()
    assert np.all(a.status.cpu().numpy() == 0)
    same = (acc == acc_ref) & (acc + rej == att_ref)
    err = rel_inf(y, y_ref)
    assert same.mean() >= count_fraction, "step counts identical for only %.4f" % same.mean()
    assert err.max() <= state_tol, "max rel state error %.3g" % err.max()
    return same.mean(), err.max()


@pytest.mark.parametrize("strict", [False, True])
def test_lorenz_rk45_golden(strict):
    g = load_golden("lorenz_rk45.npz")
    a = _run(_de().rhs_plugins.lorenz, g["y0"], 2.0, 1e-2, 1e-9, "RK45", strict=strict)
    frac, err = _compare(a, g["y_final"], g["accepted"], g["attempts"])
    assert np.max(np.abs(a[-1].t.cpu().numpy() - g["t_final"])) <= 1e-13
    print("lorenz rk45 strict=%s: counts identical %.4f, max rel err %.3g" % (strict, frac, err))
    if strict:
        # reference order of operations: differs from the reference only through pow/atan last-ulp
        assert err <= 1e-12 and frac == 1.0


@pytest.mark.parametrize("strict", [False, True])
def test_vdp_rk45_golden(strict):
    g = load_golden("vdp_rk45.npz")
    a = _run(_de().rhs_plugins.van_der_pol, g["y0"], 10.0, 1e-2, 1e-9, "RK45",
             constants=dict(mu=torch.as_tensor(g["mu"])), strict=strict)
    _compare(a, g["y_final"], g["accepted"], g["attempts"])


@pytest.mark.parametrize("strict", [False, True])
def test_lorenz_rk87_golden(strict):
    g = load_golden("lorenz_rk87.npz")
    a = _run(_de().rhs_plugins.lorenz, g["y0"], 2.0, 1e-2, 1e-9, "RK87", strict=strict)
    _compare(a, g["y_final"], g["accepted"], g["attempts"])


@pytest.mark.parametrize("name,method", [("lorenz_dopri45.npz", "Dormand-Prince"), ("lorenz_heun_euler.npz", "AHE"),
                                         ("lorenz_rk4.npz", "RK4"), ("lorenz_rk108.npz", "RK108")])
def test_other_tableaux_golden(name, method):
    g = load_golden(name)
    a = _run(_de().rhs_plugins.lorenz, g["y0"], float(g["tf"]), 1e-2, float(g["rtol"]), method)
    _compare(a, g["y_final"], g["accepted"], g["attempts"], count_fraction=1.0)


def test_edge_cases_golden():
    """Oversized dt0 (double rejections -> 3-term controller), dt0 > span clamp, tiny spans."""
    g = load_golden("edge_rk45.npz")
    de = _de()
    for i in range(g["y0"].shape[1]):
        for strict in (False, True):
            a = _run(de.rhs_plugins.lorenz, g["y0"][:, i:i + 1], float(g["tf"][i]), float(g["dt0"][i]),
                     float(g["rtol"][i]), "RK45", strict=strict)
            _compare(a, g["y_final"][:, i:i + 1], g["accepted"][i:i + 1], g["attempts"][i:i + 1], count_fraction=1.0)


def test_known_answers_single_system():
    """BASELINE config 0 (README oscillator) and the SURVEY 8c known answers, as ensemble=False."""
    de = _de()
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    hx = lambda v: np.array([float.fromhex(s) for s in v])
    two_pi = float.fromhex(ka["KA0"]["t"][0])
    for key, method in (("KA0", "RK45"), ("KA3", "RK87")):
        a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, two_pi), dt=0.01,
                         rtol=1e-9, atol=1e-9, constants=dict(k=1.0, m=1.0))
        a.method = method
        a.integrate()
        assert int(a.accepted[0]) == ka[key]["accepted"] and int(a.rejected[0]) == 0
        assert a.nfev == ka[key]["nfev"]
        # RK8(7) on this problem has rounding-noise-sized error estimates in its first steps, so dt --
        # and the state at the 1e-12 level -- depends on last-ulp details (see test_oracle_golden.py)
        assert rel_inf(a[-1].y.cpu().numpy(), hx(ka[key]["y"])) < STATE_TOL
        assert float(a[-1].t) == two_pi
        assert a.success and a.integration_status == "Integration completed successfully."
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0=np.array([1.0, 1.0, 20.0]), t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9)
    a.integrate()
    assert int(a.accepted[0]) == ka["KA1"]["accepted"] and a.nfev == ka["KA1"]["nfev"]
    assert rel_inf(a[-1].y.cpu().numpy(), hx(ka["KA1"]["y"])) < 1e-12
    for key, mu in (("KA2a", 0.1), ("KA2b", 10.0)):
        a = de.OdeSystem(de.rhs_plugins.van_der_pol, y0=np.array([2.0, 0.0]), t=(0, 10.0), dt=1e-2, rtol=1e-9,
                         atol=1e-9, constants=dict(mu=mu))
        a.integrate()
        assert int(a.accepted[0]) == ka[key]["accepted"] and a.nfev == ka[key]["nfev"]
        assert rel_inf(a[-1].y.cpu().numpy(), hx(ka[key]["y"])) < 1e-11


def test_against_oracle_fresh_inputs():
    """Seeds the fixtures never saw, at a size the oracle finishes in seconds."""
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(8192, seed=11)
    ref = orc.erk_ensemble(orc.RHS_LORENZ, y0, 0.0, 2.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"),
                           params=[10.0, 28.0, 8.0 / 3.0], threads=os.cpu_count())
    a = _run(de.rhs_plugins.lorenz, y0, 2.0, 1e-2, 1e-9, "RK45")
    frac, err = _compare(a, ref["y"], ref["accepted"], ref["accepted"] + ref["rejected"])
    print("8192 fresh Lorenz trajectories: counts identical %.5f, max rel err %.3g" % (frac, err))
    y0, mu = de.benchmarks.vdp_ensemble(4096, seed=5)
    ref = orc.erk_ensemble(orc.RHS_VDP, y0, 0.0, 10.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"),
                           params=mu[None, :], threads=os.cpu_count())
    a = _run(de.rhs_plugins.van_der_pol, y0, 10.0, 1e-2, 1e-9, "RK45", constants=dict(mu=torch.as_tensor(mu)))
    _compare(a, ref["y"], ref["accepted"], ref["accepted"] + ref["rejected"])


def test_deterministic_and_order_independent():
    """Size-independent properties at the full BASELINE size (2^20 Lorenz trajectories): two runs
    are bitwise identical although lanes pull work dynamically, and permuting the ensemble
    permutes the result (each trajectory's arithmetic is independent of its neighbours)."""
    de = _de()
    n = 1 << 20
    y0 = de.benchmarks.lorenz_ensemble(n)
    a = _run(de.rhs_plugins.lorenz, y0, 2.0, 1e-2, 1e-9, "RK45")
    b = _run(de.rhs_plugins.lorenz, y0, 2.0, 1e-2, 1e-9, "RK45")
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
    c = _run(de.rhs_plugins.lorenz, y0[:, perm.numpy()], 2.0, 1e-2, 1e-9, "RK45")
    assert torch.equal(c[-1].y.cpu(), a[-1].y.cpu()[:, perm]) and torch.equal(c.accepted.cpu(), a.accepted.cpu()[perm])
    # every trajectory landed exactly on tf, and the first 1024 are the golden prefix
    assert bool((a[-1].t == 2.0).all())
    g = load_golden("lorenz_rk45.npz")
    assert rel_inf(a[-1].y[:, :1024].cpu().numpy(), g["y_final"]).max() <= STATE_TOL
    att = (a.accepted + a.rejected).double()
    print("2^20 Lorenz: attempts mean %.1f min %d max %d" % (att.mean().item(), att.min().item(), att.max().item()))


def test_chunked_execution_matches_one_shot():
    """callback= forces one launch per accepted step; the result must be bitwise the one-shot one."""
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(256, seed=3)
    a = _run(de.rhs_plugins.lorenz, y0, 0.5, 1e-2, 1e-9, "RK45")
    b = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    calls = []
    b.integrate(callback=lambda s: calls.append(int(s.accepted.max())))
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    assert len(calls) == int(a.accepted.max()) and calls == sorted(calls)


def test_resume_with_repeated_integrate():
    """integrate(t) is resumable (reference :951-956): 0 -> 1 -> 2 visits t = 1 exactly."""
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(128, seed=4)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.integrate(1.0)
    assert bool((a[-1].t == 1.0).all()) and len(a) == 2
    a.integrate()
    assert bool((a[-1].t == 2.0).all()) and len(a) == 3
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    tab = tb.rk_arrays("RK45CKSolver")
    r1 = orc.erk_ensemble(orc.RHS_LORENZ, y0, 0.0, 1.0, 1e-2, 1e-9, 1e-9, *tab, params=[10.0, 28.0, 8.0 / 3.0])
    r2 = orc.erk_ensemble(orc.RHS_LORENZ, r1["y"], r1["t"], 2.0, r1["dt"], 1e-9, 1e-9, *tab, params=[10.0, 28.0, 8.0 / 3.0])
    assert rel_inf(a[-1].y.cpu().numpy(), r2["y"]).max() <= STATE_TOL
    assert np.array_equal(a.accepted.cpu().numpy(), r1["accepted"] + r2["accepted"])


def test_dense_output_golden():
    g = load_golden("lorenz_dense.npz")
    de = _de()
    a = _run(de.rhs_plugins.lorenz, g["y0"], 2.0, 1e-2, 1e-9, "RK45", dense_output=True, dense_capacity=512)
    assert np.array_equal((a.sol.node_count - 1).cpu().numpy(), g["n_interp"])
    for q, tq in enumerate(g["tq"]):
        got = a.sol(float(tq)).cpu().numpy()
        assert np.max(rel_inf(got, g["dense"][q])) < 1e-10
    # per-trajectory query times, a[t] indexing, end points
    tq = torch.linspace(0.1, 1.9, g["y0"].shape[1], dtype=torch.float64)
    per = a.sol(tq).cpu().numpy()
    for i in range(len(tq)):
        assert np.array_equal(per[:, i], a.sol(float(tq[i])).cpu().numpy()[:, i])
    assert torch.equal(a[2.0].y, a[-1].y) and torch.equal(a.sol(0.0), a[0].y)


def test_tolerance_failure_raises_like_reference():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(64, seed=9)
    y0[:, 5] = np.inf                                    # non-finite state: every attempt is rejected
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0, 0.1), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    with pytest.raises(de.exception_types.FailedIntegration) as ei:
        a.integrate()
    assert isinstance(ei.value.__cause__, de.exception_types.FailedToMeetTolerances)
    st = a.status.cpu().numpy()
    assert st[5] == 1 and int(a.rejected[5]) == 65 and np.all(np.delete(st, 5) == 0)
    assert not a.success and "FailedToMeetTolerances" not in a.integration_status.split("\n")[0]


def test_bad_arguments():
    de = _de()
    with pytest.raises(ValueError):
        de.OdeSystem(de.rhs_plugins.lorenz, np.ones(3), t=(0, 1, 2))
    with pytest.raises(TypeError):
        de.OdeSystem(None, np.ones(3))
    with pytest.raises(ValueError):
        de.OdeSystem(lambda t, y: y[:2], np.ones(3))
    a = de.OdeSystem(de.rhs_plugins.lorenz, np.ones(3))
    with pytest.raises(ValueError):
        a.method = "no such method"


def test_solve_ivp_wrapper():
    """Functional interface (reference differential_system.py:1202-1255): args -> constants, t_eval by repeated
    integrate(t), and scipy parity at the reference's own bar of 1e-6 (test_differential_system.py:791-868)."""
    from scipy.integrate import solve_ivp as scipy_solve_ivp
    de = _de()
    y0 = np.array([1.0, 1.0, 20.0])
    t_eval = np.linspace(0.1, 1.5, 8)
    res = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.5), y0, method="RK45", t_eval=t_eval, args=(10.0, 28.0, 8.0 / 3.0),
                       rtol=1e-10, atol=1e-10, first_step=1e-2)
    assert res.success and tuple(res.y.shape) == (3, 8) and np.allclose(res.t.cpu().numpy(), t_eval, rtol=0, atol=1e-14)
    ref = scipy_solve_ivp(lambda t, y: [10.0 * (y[1] - y[0]), y[0] * (28.0 - y[2]) - y[1], y[0] * y[1] - 8.0 / 3.0 * y[2]],
                          (0.0, 1.5), y0, method="DOP853", t_eval=t_eval, rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(res.y.cpu().numpy() - ref.y)) < 1e-6
    ens = de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 0.5), de.benchmarks.lorenz_ensemble(64), t_eval=[0.25, 0.5], ensemble=True,
                       rtol=1e-9, atol=1e-9, first_step=1e-2)
    assert tuple(ens.y.shape) == (3, 64, 2) and tuple(ens.t.shape) == (64, 2)
    with pytest.raises(ValueError):
        de.solve_ivp(de.rhs_plugins.lorenz, (0.0, 1.0), y0, t_eval=[2.0])


@pytest.mark.parametrize("strict", [False, True])
def test_time_dependent_plugins_run_in_the_fused_kernel(strict):
    """Non-autonomous right-hand sides in the persistent kernel: the stage time t + c_s h is formed per lane.
    `forced` is the reference's own test fixture (golden: forced_rk45 / forced_rk87), `duffing` the system of its
    example notebooks with per-trajectory forcing amplitude and frequency (golden: duffing_rk45)."""
    de = _de()
    for key, method in (("rk45", "RK45"), ("rk87", "RK87")):
        g = load_golden("forced_%s.npz" % key)
        a = _run(de.rhs_plugins.forced, g["y0"], 5.0, 1e-2, 1e-10, method, strict=strict)
        assert a._launches == 1                                   # one persistent launch, not the stage-level path
        _compare(a, g["y_final"], g["accepted"], g["attempts"], count_fraction=1.0 if key == "rk45" else 0.95)
    g = load_golden("duffing_rk45.npz")
    consts = dict(gamma=torch.as_tensor(g["gamma"]), omega=torch.as_tensor(g["omega"]))
    a = _run(de.rhs_plugins.duffing, g["y0"], 10.0, 1e-2, 1e-9, "RK45", constants=consts, strict=strict)
    assert a._launches == 1
    frac, err = _compare(a, g["y_final"], g["accepted"], g["attempts"])
    print("duffing strict=%s: counts identical %.4f, max rel err %.3g" % (strict, frac, err))
    # the same callable through the stage-level path (torch cos between the stage kernels) agrees
    b = de.OdeSystem(de.rhs_plugins.duffing, torch.as_tensor(g["y0"]), t=(0.0, 10.0), dt=1e-2, rtol=1e-9, atol=1e-9,
                     constants={k: v.cuda() for k, v in consts.items()}, ensemble=True, strict=strict)
    b.use_fused = False
    b.integrate()
    _compare(b, g["y_final"], g["accepted"], g["attempts"])

