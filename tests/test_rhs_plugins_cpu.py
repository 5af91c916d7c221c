"""CPU checks of the right-hand-side layer: the tagged Python callables (what the reference integrated to produce the
fixtures) agree with the oracle's C right-hand sides, DiffRHS forwards the plugin tag, benchmark inputs are
prefix-stable."""
import numpy as np
import pytest

from desolver_b200 import benchmarks as bm
from desolver_b200 import rhs_plugins as rp
from desolver_b200.differential_system import DiffRHS, rhs_prettifier
from oracle import oracle as orc


def test_python_plugins_equal_oracle_rhs():
    rng = np.random.default_rng(0)
    for _ in range(50):
        y = rng.uniform(-30, 50, 3)
        assert np.array_equal(rp.lorenz(0.0, y, 10.0, 28.0, 8.0 / 3.0), orc.rhs(orc.RHS_LORENZ, y, [10.0, 28.0, 8.0 / 3.0]))
        y2 = rng.uniform(-3, 3, 2)
        mu = float(10 ** rng.uniform(-1, 1))
        assert np.array_equal(rp.van_der_pol(0.0, y2, mu), orc.rhs(orc.RHS_VDP, y2, [mu]))
        assert np.array_equal(rp.harmonic_oscillator(0.0, y2, 2.0, 3.0), orc.rhs(orc.RHS_SHO, y2, [2.0, 3.0]))
        t = float(rng.uniform(0, 5))
        assert np.array_equal(rp.forced_oscillator(t, y2), orc.rhs(orc.RHS_FORCED, y2, None, t=t))
    # N-body: (2, nb, 3) state; same operation order (pair sums over j left to right) -> agreement to rounding
    state, m = bm.nbody_ensemble(3, bodies=8, seed=5)
    for b in range(3):
        s = np.ascontiguousarray(state[:, b])
        got = rp.nbody(0.0, s, m=m, eps2=1e-4, G=1.0)
        ref = orc.rhs(orc.RHS_NBODY, s.reshape(-1), shared=np.concatenate([[1.0, 1e-4], m])).reshape(2, 8, 3)
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    # ensemble layout (2, nb, 3, N) of the same callable == per-system calls
    ens = np.ascontiguousarray(np.moveaxis(state, 1, -1))
    full = rp.nbody(0.0, ens, m=m, eps2=1e-4, G=1.0)
    for b in range(3):
        assert np.array_equal(full[..., b], rp.nbody(0.0, np.ascontiguousarray(state[:, b]), m=m, eps2=1e-4, G=1.0))
    A, Y = bm.linear_system(16, 4)
    assert np.allclose(rp.linear(0.0, Y, A=A), A @ Y)


def test_diffrhs_counts_and_forwards_plugin_tag():
    w = DiffRHS(rp.lorenz)
    assert w.plugin_id == rp.RHS_LORENZ and w.device_plugin == "lorenz" and w.plugin_params == ("sigma", "rho", "beta")
    assert w.nfev == 0
    w(0.0, np.ones(3))
    w(0.0, np.ones(3), sigma=9.0)
    assert w.nfev == 2 and w.njev == 0

    @rhs_prettifier("dy = -y", "$\\\\dot y = -y$")
    def decay(t, y):
        return -y
    assert isinstance(decay, DiffRHS) and str(decay) == "dy = -y" and decay._repr_markdown_().startswith("$")


def test_benchmark_inputs_are_prefix_stable():
    assert np.array_equal(bm.lorenz_ensemble(4096)[:, :1000], bm.lorenz_ensemble(1000))
    y_a, mu_a = bm.vdp_ensemble(4096)
    y_b, mu_b = bm.vdp_ensemble(512)
    assert np.array_equal(y_a[:, :512], y_b) and np.array_equal(mu_a[:512], mu_b)
    assert mu_a.min() >= 0.1 and mu_a.max() <= 10.0
    A, Y0 = bm.linear_system(64, 8)
    assert np.allclose(A + A.T, -0.2 * np.eye(64))                     # near-skew with -0.1 I damping
    st, m = bm.nbody_ensemble(5, bodies=16)
    assert st.shape == (2, 5, 16, 3) and np.isclose(m.sum(), 1.0)


def test_trace_body_records_the_callables_arithmetic(tmp_path):
    """rhs_plugins.trace_body: the statement list recorded from a callable is plain C -- compiled here with gcc (no FMA
    contraction) and compared with the callable itself at random points: the same operations in the same order, so the
    results agree to the last bit for pure arithmetic, and to an ulp of libm vs numpy for sin / cos / exp / tanh."""
    import ctypes
    import subprocess
    from desolver_b200 import benchmarks as bm
    from desolver_b200 import rhs_plugins as rp

    def mixed(t, y, k, c):
        return np.stack([y[1] * 2.0 - y[2] / (1.0 + y[0] ** 2), -k * np.sin(y[0]) + np.exp(-0.5 * t) * np.tanh(y[2] / c),
                         abs(y[0]) ** 0.5 - (y[1] - 3.0) ** 3 + np.sqrt(y[2] * y[2] + 1.0)])
    cases = [(bm.rossler, 3, ("a", "b", "c"), 0), (bm.driven_pendulum, 2, ("q", "f", "w"), 1), (bm.brusselator, 2, ("a", "b"), 0),
             (mixed, 3, ("k", "c"), 1)]
    rng = np.random.default_rng(0)
    for idx, (fn, dim, params, ulps) in enumerate(cases):
        body, timed = rp.trace_body(fn, dim, params)
        assert timed == (fn in (bm.driven_pendulum, mixed))
        src = tmp_path / ("body%d.c" % idx)
        src.write_text("#include <math.h>\nvoid f(double t, const double *y, const double *p, double *dy)\n{\n%s\n}\n" % body)
        so = tmp_path / ("body%d.so" % idx)
        subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(src), "-lm"], check=True)
        f = ctypes.CDLL(str(so)).f
        f.argtypes = [ctypes.c_double] + [np.ctypeslib.ndpointer(dtype=np.float64)] * 3
        for _ in range(200):
            t, y, p = float(rng.uniform(0, 3)), rng.uniform(-2, 2, dim), rng.uniform(0.3, 3.0, len(params))
            want = np.asarray(fn(t, y, **dict(zip(params, p))), dtype=np.float64)
            got = np.empty(dim)
            f(t, y, p, got)
            # pure arithmetic: bit for bit; with libm functions in the terms: a few ulps of the TERMS (values here are O(10))
            assert np.all(np.abs(got - want) <= ulps * 2e-14), (fn.__name__, got, want)
    with pytest.raises(NotImplementedError):
        rp.trace_body(lambda t, y: np.stack([np.sign(y[0]), y[1]]), 2)        # no device counterpart registered
    with pytest.raises(ValueError):
        rp.trace_body(lambda t, y: np.stack([y[0]]), 2)
