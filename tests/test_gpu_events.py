"""GPU tests of device event detection (dsb_erk_run_args.n_events): hyperplane events located on the step
interpolant inside the fused kernel, against tests/golden/events_*.npz -- event lists produced by the unmodified
reference, one OdeSystem per trajectory (oracle/make_golden_events.py).

The reference reports a bracketed root only when |g(root)| <= 4 eps absolute (utilities/optimizer.py:194-197, 316),
so last-ulp noise makes it miss events (on the Lorenz fixtures it reports 1.2 of the 2.4 sign changes per trajectory
that its own step history contains); the device reports every bracketed root.  Hence: every reference event must be
found by the device (same event, same time, same state), and the device's counts must equal the number of sign
changes in the reference's step history (fixture key `crossings`)."""
import numpy as np
import pytest

from conftest import load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

T_TOL = 1e-9       # event time: the root of the same cubic, found by a different root finder
Y_TOL = 1e-8


def _event_set(name):
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_golden_events", os.path.join(ROOT, "oracle", "make_golden_events.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.event_set(name)


def _run(name, events=True, use_fused=True):
    import desolver_b200 as de
    g = load_golden("events_%s.npz" % name)
    lorenz = name.startswith("lorenz")
    rhs = de.rhs_plugins.lorenz if lorenz else de.rhs_plugins.van_der_pol
    consts = {} if lorenz else dict(mu=torch.as_tensor(g["mu"]))
    a = de.OdeSystem(rhs, torch.as_tensor(np.ascontiguousarray(g["y0"])), t=(0.0, 2.0 if lorenz else 10.0), dt=1e-2, rtol=1e-9,
                     atol=1e-9, constants=consts, ensemble=True)
    a.method = "RK45"
    a.use_fused = use_fused        # False: the stage-level path (user RHS between the stage kernels) with its event kernels
    evs = _event_set(name)
    a.integrate(events=evs if events else None)
    torch.cuda.synchronize()
    return a, g, evs


def _check_events(a, g):
    log = a.events
    assert not log.overflowed
    cnt = log.count.cpu().numpy()
    t = log.t.cpu().numpy()
    y = log.y.cpu().numpy()
    idx = log.index.cpu().numpy()
    n = cnt.shape[0]
    missing, same_count, worst_t, worst_y = 0, 0, 0.0, 0.0
    for i in range(n):
        ours = [(t[i, k], y[:, i, k], idx[i, k]) for k in range(cnt[i])]
        assert all(ours[k][0] <= ours[k + 1][0] for k in range(len(ours) - 1)), "events out of time order"
        same_count += int(cnt[i] == g["count"][i])
        for k in range(min(int(g["count"][i]), g["records"].shape[1])):
            rt, ry, re = g["records"][i, k, 0], g["records"][i, k, 1:-1], int(g["records"][i, k, -1])
            cand = [o for o in ours if o[2] == re and abs(o[0] - rt) <= T_TOL * (1 + abs(rt))]
            if not cand:
                missing += 1
                continue
            worst_t = max(worst_t, abs(cand[0][0] - rt))
            worst_y = max(worst_y, np.max(np.abs(cand[0][1] - ry)) / max(1.0, np.max(np.abs(ry))))
    print("events: %d trajectories, same count as the reference's list %.3f, reference events missing %d, worst dt %.2e, "
          "worst dy %.2e" % (n, same_count / n, missing, worst_t, worst_y))
    assert missing == 0
    assert worst_y <= Y_TOL
    return cnt, idx


@pytest.mark.parametrize("use_fused", [True, False])
@pytest.mark.parametrize("name", ["lorenz_sections", "vdp_sections", "lorenz_zmax"])
def test_non_terminal_events_match_reference_and_do_not_perturb_the_integration(name, use_fused):
    a, g, evs = _run(name, use_fused=use_fused)
    cnt, idx = _check_events(a, g)
    assert cnt.sum() > 0 and a.success and "completed" in a.integration_status
    # one event per sign change of g over an accepted step, per event function
    per_event = np.stack([[(idx[i, :cnt[i]] == e).sum() for e in range(len(evs))] for i in range(cnt.shape[0])])
    same = np.all(per_event == g["crossings"], axis=1)
    print("event counts equal the sign changes of the reference's step history for %.3f of the trajectories" % same.mean())
    assert same.mean() >= 0.97
    assert np.all(a.status.cpu().numpy() == 0)
    # recorded states lie on their hyperplanes
    log = a.events
    import desolver_b200 as de
    for i in range(0, cnt.shape[0], 7):
        for st in log.for_trajectory(i):
            if st.event.requires_dstate:
                # extremum of z along the INTERPOLANT; the true derivative there vanishes to the interpolation error
                dy = de.rhs_plugins.lorenz(float(st.t), st.y.numpy())
                assert abs(float(st.event(float(st.t), st.y.numpy(), dy))) <= 1e-2
            else:
                assert abs(float(st.event(float(st.t), st.y.numpy()))) <= 1e-9
    # events are observers: the integration itself is bit-identical to a run without them
    b, _, _ = _run(name, events=False, use_fused=use_fused)
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    assert rel_inf(a[-1].y.cpu().numpy(), g["y_final"]).max() <= 1e-9
    assert np.mean(a.accepted.cpu().numpy() == g["accepted"]) >= 0.97


@pytest.mark.parametrize("use_fused", [True, False])
@pytest.mark.parametrize("name", ["lorenz_terminal", "vdp_terminal"])
def test_terminal_events_stop_the_trajectory_at_the_root(name, use_fused):
    import desolver_b200 as de
    from desolver_b200.backend import cuda_backend as cb
    a, g, evs = _run(name, use_fused=use_fused)
    _check_events(a, g)
    st = a.status.cpu().numpy()
    tfin = a[-1].t.cpu().numpy()
    yfin = a[-1].y.cpu().numpy()
    tf = 2.0 if name.startswith("lorenz") else 10.0
    dev_term = st == cb.TRAJ_EVENT
    ref_term = g["status"] == 2
    # the reference can miss a terminal crossing (see the module docstring) and then runs on; never the other way round
    assert np.all(dev_term[ref_term])
    assert "terminated upon finding a triggered event" in a.integration_status
    same_stop = ref_term & (np.abs(tfin - g["t_final"]) <= T_TOL * 10)
    print("terminated: device %d, reference %d, same stopping time %d" % (dev_term.sum(), ref_term.sum(), same_stop.sum()))
    assert same_stop.sum() >= 0.5 * ref_term.sum()
    assert rel_inf(yfin[:, same_stop], g["y_final"][:, same_stop]).max() <= 1e-8
    assert np.mean(a.accepted.cpu().numpy()[same_stop] == g["accepted"][same_stop]) >= 0.95
    # every stopped trajectory: the terminal event is the last recorded one, the final time is its root, and the state
    # reached by re-integrating to the root sits on the hyperplane to the accuracy of the cubic interpolant
    term_ev = [e for e in evs if e.is_terminal][0]
    for i in np.nonzero(dev_term)[0][::3]:
        lst = a.events.for_trajectory(int(i))
        assert lst and lst[-1].event is term_ev and abs(float(lst[-1].t) - tfin[i]) <= 1e-12
        assert abs(float(term_ev(tfin[i], yfin[:, i]))) <= 1e-5     # cubic Hermite interpolation error, O(h^4)
        assert sum(1 for s_ in lst if s_.event is term_ev) == 1
    # untouched trajectories ran to tf
    assert np.all(st[~dev_term] == 0) and np.all(np.abs(tfin[~dev_term] - tf) <= 1e-12)


def test_single_system_event_list_has_the_reference_shape():
    import desolver_b200 as de
    ev = de.events.component_crossing(0, 0.0, dim=2, direction=-1)
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.01, rtol=1e-9, atol=1e-9)
    a.integrate(events=ev)
    lst = a.events
    assert isinstance(lst, list) and len(lst) == 1 and lst[0].event is ev
    assert abs(float(lst[0].t) - np.pi / 2) <= 1e-8 and abs(float(lst[0].y[0])) <= 1e-12
    # plain Python callables of affine form are recognised by probing and run on the device (the reference's own
    # event tests use exactly these: a time mark and a level of one component, tests/test_event_detection.py:22-29,83-91)
    def time_event(t, y, **kwargs):
        return np.asarray(t - np.pi / 8)
    time_event.is_terminal = True
    time_event.direction = 0

    def first_y_event(t, y, **kwargs):
        return y[0] - np.cos(np.pi / 16)
    a.reset()
    a.integrate(events=[time_event, first_y_event])
    assert a.integration_status == "Integration terminated upon finding a triggered event."
    assert abs(float(a[-1].t) - np.pi / 8) <= 1e-12
    lst = a.events
    assert [st.event for st in lst] == [first_y_event, time_event]
    assert abs(float(lst[0].t) - np.pi / 16) <= 1e-5 and abs(float(lst[1].t) - np.pi / 8) <= 1e-12
    # non-affine events are not refused any more (round 2): they go through the event-search kernels
    a.reset()
    a.integrate(events=lambda t, y: (t - 0.1) * (t - 0.2))
    assert [round(float(st.t), 12) for st in a.events] == [0.1, 0.2]


def test_events_with_an_untagged_callable_and_a_13_stage_method():
    """Stage-level path: a user torch callable (no plugin) with RK8(7), one recorded and one terminal event, against the
    closed-form solution of the forced oscillator y'' + y = t: y = t + (y0 - 0) cos t + (v0 - 1) sin t."""
    import desolver_b200 as de
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.5, 1.5, size=(2, 300))
    a = de.OdeSystem(lambda t, y: torch.stack([y[1], -y[0] + t]), torch.as_tensor(y0), t=(0.0, 6.0), dt=1e-2, rtol=1e-10,
                     atol=1e-10, ensemble=True)
    a.method = "RK87"
    level = de.events.HyperplaneEvent([1.0, 0.0], c=0.0, wt=-1.0, direction=0)         # y - t = 0: a cos t + b sin t = 0
    stop = de.events.component_crossing(1, 1.0, dim=2, direction=-1, is_terminal=True)  # v falls through 1: -a sin t + b cos t = 0
    a.integrate(events=[level, stop])
    log = a.events
    A, B = y0[0], y0[1] - 1.0
    t_level = np.mod(np.arctan2(-A, B), np.pi)                   # first root of A cos t + B sin t
    t_stop_all = np.mod(np.arctan2(B, A), np.pi)                 # roots of -A sin t + B cos t (every pi)
    tfin = a[-1].t.cpu().numpy()
    st = a.status.cpu().numpy()
    # a falling crossing of v = 1 comes once per period 2 pi > 6: most, not all, trajectories meet one
    assert set(np.unique(st)) <= {0, 3} and (st == 3).mean() > 0.5 and np.all(np.abs(tfin[st == 0] - 6.0) <= 1e-12)
    for i in np.nonzero(st == 3)[0][::7]:
        lst = log.for_trajectory(int(i))
        assert lst[-1].event is stop and abs(float(lst[-1].t) - tfin[i]) <= 1e-12
        # the terminal root is one of the stationary points of v with v' < 0 there
        k = np.round((tfin[i] - t_stop_all[i]) / np.pi)
        assert abs(tfin[i] - (t_stop_all[i] + k * np.pi)) <= 2e-5      # cubic Hermite interpolation over RK8(7)-sized steps
        levels = [float(st.t) for st in lst if st.event is level]
        for tl in levels:
            kk = np.round((tl - t_level[i]) / np.pi)
            assert abs(tl - (t_level[i] + kk * np.pi)) <= 2e-5
    # final state = exact solution at the stopping time, to the integration tolerance
    exact = tfin + A * np.cos(tfin) + B * np.sin(tfin)
    assert np.max(np.abs(a[-1].y[0].cpu().numpy() - exact)) <= 1e-7


def test_events_with_a_symplectic_method():
    """Splitting scheme + events (the reference's event tests run ABAs5o6H too, tests/test_event_detection.py): harmonic
    oscillators q'' = -q with different amplitudes, a recorded section q = 0 and a terminal time mark."""
    import desolver_b200 as de
    amp = np.linspace(0.5, 2.0, 64)
    y0 = np.stack([amp, np.zeros(64)])
    a = de.OdeSystem(lambda t, y: torch.stack([y[1], -y[0]]), torch.as_tensor(y0), t=(0.0, 10.0), dt=0.01, ensemble=True)
    a.method = "ABAS5O6H"
    section = de.events.component_crossing(0, 0.0, dim=2, direction=0)

    def time_mark(t, y, **kwargs):              # a plain Python callable of affine form, terminal
        return t - 7.3
    time_mark.is_terminal = True
    a.integrate(events=[section, time_mark])
    assert a.integration_status == "Integration terminated upon finding a triggered event."
    assert np.all(a.status.cpu().numpy() == 3) and np.max(np.abs(a[-1].t.cpu().numpy() - 7.3)) <= 1e-12
    log = a.events
    for i in range(0, 64, 9):
        lst = log.for_trajectory(i)
        zeros = [float(st.t) for st in lst if st.event is section]
        assert lst[-1].event is time_mark and abs(float(lst[-1].t) - 7.3) <= 1e-12
        assert np.allclose(zeros, [np.pi / 2, 3 * np.pi / 2], rtol=0, atol=1e-8) and len(zeros) == 2
    yf = a[-1].y.cpu().numpy()
    assert np.max(np.abs(yf[0] - amp * np.cos(7.3))) <= 1e-9 and np.max(np.abs(yf[1] + amp * np.sin(7.3))) <= 1e-9


def test_event_edge_cases():
    """Record-store overflow, strict fused vs stage-level path (bit-equal event lists), a terminal event inside the very
    first step of a single system, and events on the second leg of a continued integration."""
    import desolver_b200 as de
    cc = de.events.component_crossing
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(500, seed=4)).cuda()
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 5.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.event_capacity = 2
    a.integrate(events=[cc(0, 0.0, dim=3), cc(1, 0.0, dim=3)])
    log = a.events
    assert log.overflowed and int(log.count.max()) > 2 and len(log.for_trajectory(int(log.count.argmax()))) == 2
    res = []
    for fused in (True, False):
        b = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, strict=True)
        b.use_fused = fused
        b.integrate(events=[cc(0, 0.0, dim=3)])
        res.append((b.events.count.clone(), b.events.t.clone(), b[-1].y.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
    valid = torch.arange(res[0][1].shape[1], device="cuda")[None, :] < res[0][0][:, None]
    assert float((res[0][1] - res[1][1])[valid].abs().max()) <= 1e-15
    s = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, 1.0), dt=0.5, rtol=1e-9, atol=1e-9)
    s.integrate(events=cc(0, 0.99999, dim=2, direction=-1, is_terminal=True))
    assert "terminated" in s.integration_status and abs(float(s[-1].t) - np.arccos(0.99999)) <= 1e-5      # root of the cubic interpolant of a large first step
    assert len(s.events) == 1 and float(s.events[0].t) == float(s[-1].t)
    c = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    c.integrate(1.0)
    c.integrate(2.0, events=[cc(0, 0.0, dim=3)])
    d = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    d.integrate(1.0)
    d.integrate(2.0)
    assert torch.equal(c[-1].y, d[-1].y) and float(c.events.t[c.events.count > 0][:, 0].min()) >= 1.0
