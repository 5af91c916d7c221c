"""GPU parity tests of the explicit symplectic splitting integrators (K4): the fused N-body kernel and the
stage-level kernels, against the reference fixtures and the CPU oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _nbody_system(state, m, tf, dt, method, strict=False, use_fused=True, eps2=1e-4):
    """state (2, B, nb, 3) reference layout -> ensemble layout (2, nb, 3, B)."""
    de = _de()
    y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1)))
    a = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, tf), dt=dt, constants=dict(m=torch.as_tensor(m), eps2=eps2, G=1.0),
                     ensemble=True, strict=strict)
    a.method = method
    a.use_fused = use_fused
    a.integrate()
    return a, np.moveaxis(a[-1].y.cpu().numpy(), -1, 1)


@pytest.mark.parametrize("method", ["BABS9O7H", "ABAS5O6H"])
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("use_fused", [True, False])
def test_nbody_golden(method, strict, use_fused):
    g = load_golden("nbody_symplectic.npz")
    a, y = _nbody_system(g["state0"], g["m"], float(g["tf"]), float(g["dt"]), method, strict, use_fused)
    ref = g[method + ".y"]
    steps = len(g[method + ".t"]) - 1
    assert np.all(a.accepted.cpu().numpy() == steps) and np.all(a.status.cpu().numpy() == 0)
    assert np.all(a[-1].t.cpu().numpy() == g[method + ".t"][-1])          # truncated last step lands on tf
    B = ref.shape[1]
    err = max(rel_inf(y[:, b].reshape(-1), ref[:, b].reshape(-1), axis=0) for b in range(B))
    assert err <= (1e-13 if strict else 1e-11), err
    assert a.nfev == B * (2 + a.integrator.stages * steps)                  # reference: 1 + S*steps + 1 per system


def test_nbody_against_oracle_64_bodies():
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    de = _de()
    state, m = de.benchmarks.nbody_ensemble(48, bodies=64, seed=4)
    a, y = _nbody_system(state, m, 0.1, 0.01, "BABS9O7H")
    B = state.shape[1]
    y0 = np.ascontiguousarray(np.moveaxis(state, 1, -1).reshape(-1, B))
    shared = np.concatenate([[1.0, 1e-4], m])
    ref = orc.symplectic_ensemble(orc.RHS_NBODY, y0, 0.0, 0.1, 0.01, tb.symplectic_array("BABs9o7HSolver")[0],
                                  shared=shared, threads=os.cpu_count())
    got = a[-1].y.reshape(-1, B).cpu().numpy()
    assert np.all(a.accepted.cpu().numpy() == ref["steps"]) and np.all(ref["steps"] == 10)
    assert rel_inf(got, ref["y"]).max() <= 1e-11
    # size-independent properties: total momentum is conserved, and the scheme is time-reversible
    mom0 = (m[None, :, None] * state[1]).sum(1)
    mom1 = (m[None, :, None] * y[1]).sum(1)
    assert np.max(np.abs(mom1 - mom0)) < 1e-14
    # time reversibility (size-independent property): flip the velocities, integrate the same 8 dyadic steps
    # forward again, flip back -> the initial state, to roundoff.  (Integrating towards a SMALLER time is not
    # usable for this: the reference's final-step test |dt + t| > |tf| (differential_system.py:997) is true
    # at once for decreasing positive times, so it takes one single step; the engine mirrors that.)
    dt = 2.0 ** -7
    f, yf = _nbody_system(state, m, 8 * dt, dt, "BABS9O7H")
    assert np.all(f.accepted.cpu().numpy() == 8)
    flipped = yf.copy()
    flipped[1] *= -1.0
    b, yb = _nbody_system(flipped, m, 8 * dt, dt, "BABS9O7H")
    yb[1] *= -1.0
    diff = np.abs(yb - state)
    assert np.median(diff) < 1e-13 and np.max(diff) < 1e-8       # roundoff, amplified only by close encounters


def test_backward_time_mirrors_reference_single_step():
    """Reference semantics for tf < t0 with positive times: one truncated step (see the note above)."""
    de = _de()
    a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(1.0, 0.5), dt=0.05, strict=True)
    a.method = "BABS9O7H"
    a.integrate()
    assert int(a.accepted[0]) == 1 and float(a[-1].t) == 0.5 and float(a.dt) == -0.05


def test_known_answers_sho_symplectic():
    """KA4 / KA5: README oscillator, dt = 0.05, 126 steps, last one truncated; stage-level
    kernels around the Python callable (eager first step, then CUDA-graph replay)."""
    de = _de()
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for key, method in (("KA4", "BABS9O7H"), ("KA5", "ABAS5O6H")):
        two_pi = float.fromhex(ka[key]["t"][0])
        want = np.array([float.fromhex(s) for s in ka[key]["y"]])
        for strict in (True, False):
            for graphs in (True, False):
                a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, two_pi), dt=0.05, strict=strict)
                a.method = method
                a.use_graphs = graphs
                a.integrate()
                assert int(a.accepted[0]) == ka[key]["steps"] == 126
                assert float(a[-1].t) == two_pi
                assert a.nfev == ka[key]["nfev"]
                y = a[-1].y.cpu().numpy()
                if strict:
                    assert np.array_equal(y, want)                              # +,* only: bit-exact
                else:
                    assert np.max(np.abs(y - want)) < 1e-14


def test_odd_leading_axis_kick_mask():
    """Kick variables are rows dim0 // 2 .. of axis 0 (reference integrator_types.py:359-362): for a 3-row state that is
    1 drift row and 2 kick rows.  Engine (stage-level kernels) vs oracle with the same split; the reference
    semantics (rhs evaluated on the whole state, unused half masked by 0) make this well defined for any RHS."""
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    de = _de()
    rng = np.random.default_rng(2)
    A = rng.standard_normal((3, 3)) * 0.3
    y0 = rng.standard_normal((3, 16))
    a = de.OdeSystem(de.rhs_plugins.linear, torch.as_tensor(y0), t=(0.0, 0.32), dt=0.05, constants=dict(A=torch.as_tensor(A).cuda()),
                     ensemble=True, strict=True)
    a.method = "ABAS5O6H"
    a.integrate()
    ref = orc.symplectic_ensemble(orc.RHS_LINEAR, y0, 0.0, 0.32, 0.05, tb.symplectic_array("ABAs5o6HSolver")[0], shared=A,
                                  drift_count=1)
    assert np.array_equal(a.accepted.cpu().numpy(), ref["steps"])
    assert rel_inf(a[-1].y.cpu().numpy(), ref["y"]).max() < 1e-13
    wrong = orc.symplectic_ensemble(orc.RHS_LINEAR, y0, 0.0, 0.32, 0.05, tb.symplectic_array("ABAs5o6HSolver")[0], shared=A,
                                    drift_count=2)
    assert rel_inf(a[-1].y.cpu().numpy(), wrong["y"]).max() > 1e-6


def test_symplectic_dense_output():
    """Dense output of the splitting schemes (stage-level path writes a node (t, y, rhs(t, y)) per step).  KA7: inside
    the first step the interpolant equals the reference's bit for bit; later the reference uses a stale end slope
    (integrators/integrator_types.py:385-386) while the nodes here carry their own, so accuracy is checked against the
    analytic solution instead."""
    import json
    de = _de()
    ka = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))["KA7"]
    a = de.OdeSystem(lambda t, y: torch.stack([y[1], -y[0]]), np.array([1.0, 0.0]), t=(0, 1.0), dt=0.05, dense_output=True, strict=True)
    a.method = "ABAS5O6H"
    a.integrate()
    assert int(a.accepted[0]) == ka["steps"]
    got = a.sol(0.02).cpu().numpy()
    assert [float.hex(float(v)) for v in got] == ka["sol_0p02"]
    assert [float.hex(float(v)) for v in a.sol(0.05).cpu().numpy()] == ka["sol_0p05"]
    for tq in (0.13, 0.53, 0.999):
        y = a.sol(tq).cpu().numpy()
        assert abs(y[0] - np.cos(tq)) < 2e-7 and abs(y[1] + np.sin(tq)) < 2e-7          # cubic Hermite, h = 0.05
    assert np.array_equal(a.sol(1.0).cpu().numpy(), a[-1].y.cpu().numpy())
    # ensemble: N-body systems, the interpolant at the end time is the final state, at t = 0 the initial one
    state, m = de.benchmarks.nbody_ensemble(12, bodies=16, seed=8)
    y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1)))
    b = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, 0.05), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0),
                     ensemble=True, dense_output=True)
    b.method = "BABS9O7H"
    b.integrate()
    assert int(b.sol.node_count.min()) == 6 and torch.equal(b.sol(0.05), b[-1].y) and torch.equal(b.sol(0.0).cpu(), y0)
    mid = b.sol(0.025)
    assert float((mid - b[-1].y).abs().max()) > 0 and bool(torch.isfinite(mid).all())
