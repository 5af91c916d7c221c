"""CPU tests of the event objects (no GPU): HyperplaneEvent is a valid event callable of the reference's protocol
on numpy arrays and torch tensors, and the committed event fixtures are well formed."""
import numpy as np

from conftest import load_golden


def test_hyperplane_event_protocol():
    import torch
    from desolver_b200.events import HyperplaneEvent, component_crossing
    ev = HyperplaneEvent([1.0, 0.0, -2.0], c=0.5, wt=0.25, direction=-3, is_terminal=True)
    y = np.array([1.0, 7.0, 0.125])
    assert ev(2.0, y) == 1.0 - 0.25 + 0.25 * 2.0 - 0.5
    assert ev.direction == -1 and ev.is_terminal and ev.flags == (2 | 4)
    assert np.array_equal(ev.coefficients(3), [1.0, 0.0, -2.0, 0.25, 0.5])
    ens = np.stack([y, 2 * y], axis=-1)
    assert np.array_equal(ev(2.0, ens), [ev(2.0, y), ev(2.0, 2 * y)])
    assert torch.equal(ev(2.0, torch.as_tensor(ens)), torch.as_tensor(ev(2.0, ens)))
    c = component_crossing(1, 3.0, dim=2, direction=1)
    assert c(0.0, np.array([9.0, 4.0])) == 1.0 and c.flags == 1 and not c.is_terminal
    # multi-axis state (2, 2) flattened to 4 components
    m = component_crossing(3, 0.0, dim=4)
    assert m(0.0, np.arange(4.0).reshape(2, 2)) == 3.0


def test_event_fixtures_are_well_formed():
    for name in ("lorenz_sections", "lorenz_terminal", "vdp_sections", "vdp_terminal"):
        g = load_golden("events_%s.npz" % name)
        n = g["y0"].shape[1]
        assert g["records"].shape[0] == n and g["count"].shape == (n,)
        for i in range(n):
            k = int(g["count"][i])
            t = g["records"][i, :k, 0]
            assert np.all(np.isfinite(t)) and np.all(np.diff(t) >= 0)
        if name.endswith("terminal"):
            term = g["status"] == 2
            assert term.any()
            # a terminated trajectory ends at its last event
            last = np.array([g["records"][i, g["count"][i] - 1, 0] for i in np.nonzero(term)[0]])
            assert np.max(np.abs(last - g["t_final"][term])) <= 1e-9


def test_affine_callables_are_recognised_and_others_refused():
    import pytest
    import torch
    from desolver_b200.events import as_hyperplane

    def time_event(t, y, **kwargs):
        return np.asarray(t - np.pi / 8)
    time_event.is_terminal = True
    h = as_hyperplane(time_event, 2)
    assert h.source is time_event and h.is_terminal and h.wt == 1.0 and abs(h.c - np.pi / 8) < 1e-16 and not h.w.any()
    h = as_hyperplane(lambda t, y, k=2.0: y[1] + k * y[0] - 0.5, 2, dict(k=3.0))
    assert np.allclose(h.w, [3.0, 1.0]) and h.wt == 0.0 and h.c == 0.5
    h = as_hyperplane(lambda t, y: torch.as_tensor(y[0]).to(torch.float64) - 1.0, 3)      # torch-only callable
    assert np.array_equal(h.w, [1.0, 0.0, 0.0]) and h.c == 1.0
    for bad in (lambda t, y: (t - 1.0) * (t - 2.0), lambda t, y: y[0] * y[1], lambda t, y: np.sin(y[0])):
        with pytest.raises(NotImplementedError):
            as_hyperplane(bad, 2)
    with pytest.raises(NotImplementedError):                  # depends on a per-trajectory constant
        as_hyperplane(lambda t, y, mu=None: y[0] - mu, 2, dict(mu=np.linspace(0.0, 1.0, 5)))
    ev = lambda t, y, dy: 2.0 * dy[1] - y[0] + 0.5 * t - 3.0      # the reference's requires_dstate protocol
    ev.requires_dstate = True
    h = as_hyperplane(ev, 2)
    assert h.requires_dstate and np.array_equal(h.w_dy, [0.0, 2.0]) and np.array_equal(h.w, [-1.0, 0.0])
    assert h.wt == 0.5 and h.c == 3.0
    assert h(1.0, np.array([1.0, 0.0]), np.array([0.0, 4.0])) == ev(1.0, np.array([1.0, 0.0]), np.array([0.0, 4.0]))
    bad = lambda t, y, dy: dy[0] * y[0]
    bad.requires_dstate = True
    with pytest.raises(NotImplementedError):
        as_hyperplane(bad, 2)


def test_callable_event_wrapper_and_the_affine_probe():
    """Events that are not affine fall through to CallableEvent (evaluated between the event-search kernels); affine
    callables are recognised by probing and keep the closed-form device path."""
    import pytest
    import torch
    from desolver_b200.events import CallableEvent, HyperplaneEvent, as_hyperplane
    from oracle import event_cases as ec
    hp = as_hyperplane(lambda t, y: 2.0 * y[1] - t + 0.5, 3)
    assert isinstance(hp, HyperplaneEvent) and np.allclose(hp.coefficients(3), [0.0, 2.0, 0.0, -1.0, -0.5])
    for fn in (ec.lorenz_sphere, ec.lorenz_product, ec.time_cubic):
        with pytest.raises(NotImplementedError):
            as_hyperplane(fn, 3)
    with pytest.raises(NotImplementedError):                     # depends on a per-trajectory constant
        as_hyperplane(ec.vdp_energy_mu, 2, dict(mu=torch.linspace(0.1, 1.0, 5)))
    ev = CallableEvent(ec.lorenz_sphere_terminal)
    assert ev.is_terminal and ev.direction == 1 and ev.flags == (1 | 4) and not ev.requires_dstate
    assert np.array_equal(ev.coefficients(3), np.zeros(5)) and ev.signature[0] == "callable"
    y = torch.tensor([[3.0, 30.0], [4.0, 0.0], [25.0, 25.0]], dtype=torch.float64)
    assert torch.equal(ev(0.0, y), torch.tensor([25.0 - 484.0, 900.0 - 484.0], dtype=torch.float64))
    dy = CallableEvent(ec.radius_maximum)
    assert dy.requires_dstate and dy.flags == (2 | 4)


def test_nonaffine_event_fixtures_are_well_formed():
    from oracle import event_cases as ec
    for name, (rhs_name, evs, tf, tol, method) in ec.CASES.items():
        g = load_golden("events2_%s.npz" % name)
        n = g["y0"].shape[1]
        assert g["records"].shape[0] == n and g["crossings"].shape == (n, len(evs))
        # the reference reports a SUBSET of the sign changes it steps over (absolute root tolerance, see the test module)
        assert np.all(g["count"] <= g["crossings"].sum(1) + 1)
