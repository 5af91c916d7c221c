"""Synthetic inputs of the BASELINE.json configs, seeded and prefix-stable:
the first M trajectories are the same for every ensemble size N >= M, so the golden fixtures
(first 1024 / 512 trajectories) are literally a prefix of the 2^20 / 2^24 benchmark ensembles."""
import numpy as np

LORENZ = dict(sigma=10.0, rho=28.0, beta=8.0 / 3.0, t0=0.0, tf=2.0, dt0=1e-2, rtol=1e-9, atol=1e-9)
VDP = dict(t0=0.0, tf=10.0, dt0=1e-2, rtol=1e-9, atol=1e-9)
NBODY = dict(bodies=64, G=1.0, eps2=1e-4, t0=0.0, tf=0.1, dt=0.01)
LINEAR = dict(t0=0.0, tf=1.0, dt0=1e-2, rtol=1e-9, atol=1e-9)


def lorenz_ensemble(n, seed=0):
    """y0 (3, n) ~ U([-20,20] x [-25,25] x [0,50])."""
    rng = np.random.default_rng(seed)
    y0 = rng.uniform([-20.0, -25.0, 0.0], [20.0, 25.0, 50.0], size=(n, 3))
    return np.ascontiguousarray(y0.T)


def vdp_ensemble(n, seed=0):
    """(y0 (2, n) ~ U([-2,2]^2), mu (n,) with log10 mu ~ U(-1, 1))."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(size=(n, 3))
    mu = 10.0 ** (2.0 * u[:, 0] - 1.0)
    y0 = 4.0 * u[:, 1:] - 2.0
    return np.ascontiguousarray(y0.T), np.ascontiguousarray(mu)


def linear_system(n, columns, seed=0):
    """(A (n, n) near-skew with -0.1 I damping, Y0 (n, columns) ~ N(0,1))."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n))
    a = (g - g.T) / np.sqrt(2.0 * n) - 0.1 * np.eye(n)
    y0 = rng.standard_normal((n, columns))
    return np.ascontiguousarray(a), np.ascontiguousarray(y0)


def nbody_ensemble(systems, bodies=64, seed=0):
    """state (2, systems, bodies, 3) = (q ~ N(0,1); v ~ 0.1 N(0,1)), masses (bodies,) = 1/bodies."""
    rng = np.random.default_rng(seed)
    qv = rng.standard_normal((systems, 2, bodies, 3))
    qv[:, 1] *= 0.1
    state = np.ascontiguousarray(np.moveaxis(qv, 1, 0))
    return state, np.full(bodies, 1.0 / bodies)


# ---- an example of a caller-defined fused device plugin (rhs_plugins.register_device_rhs): the Roessler system ----
def rossler(t, y, a=0.2, b=0.2, c=5.7):
    """Roessler attractor, written once for numpy / torch (any trailing ensemble axis) ..."""
    parts = [-y[1] - y[2], y[0] + a * y[1], b + y[2] * (y[0] - c)]
    if type(y).__module__.split(".")[0] == "torch":
        import torch
        return torch.stack(parts)
    return np.stack(parts)


# ... and once as the CUDA body of the fused kernel's right-hand side (same operations, left to right)
ROSSLER_BODY = """
dy[0] = -y[1] - y[2];
dy[1] = y[0] + p[0] * y[1];
dy[2] = p[1] + y[2] * (y[0] - p[2]);
"""


def rossler_plugin():
    """`rossler` tagged with its fused device plugin (built on first use, about a minute; `__graft_entry__.build()`
    prebuilds it in-tree)."""
    from . import rhs_plugins
    return rhs_plugins.register_device_rhs(rossler, ROSSLER_BODY, dim=3, params=("a", "b", "c"),
                                           defaults=dict(a=0.2, b=0.2, c=5.7), name="rossler")


# ---- a second example: time-dependent, two components, transcendental functions (the damped driven pendulum) ----
def driven_pendulum(t, y, q=0.5, f=1.2, w=2.0 / 3.0):
    if type(y).__module__.split(".")[0] == "torch":
        import torch
        return torch.stack([y[1], -torch.sin(y[0]) - q * y[1] + f * torch.cos(w * t)])
    return np.stack([y[1], -np.sin(y[0]) - q * y[1] + f * np.cos(w * t)])


DRIVEN_PENDULUM_BODY = """
dy[0] = y[1];
dy[1] = -sin(y[0]) - p[0] * y[1] + p[1] * cos(p[2] * t);
"""


def driven_pendulum_plugin():
    from . import rhs_plugins
    return rhs_plugins.register_device_rhs(driven_pendulum, DRIVEN_PENDULUM_BODY, dim=2, params=("q", "f", "w"),
                                           defaults=dict(q=0.5, f=1.2, w=2.0 / 3.0), time_dependent=True, name="driven_pendulum")


# ---- a third example, WITHOUT a hand-written body: `register_device_rhs(fn, dim=...)` evaluates the callable once on
# symbolic scalars (rhs_plugins.trace_body) and compiles what it recorded -- the Brusselator ----
def brusselator(t, y, a=1.0, b=3.0):
    x, v = y[0], y[1]
    parts = [a + x ** 2 * v - (b + 1.0) * x, b * x - x ** 2 * v]
    if type(x).__module__.split(".")[0] == "torch":
        import torch
        return torch.stack(parts)
    return np.stack(parts) if isinstance(x, np.ndarray) or np.isscalar(x) else parts


def brusselator_plugin():
    from . import rhs_plugins
    return rhs_plugins.register_device_rhs(brusselator, dim=2, params=("a", "b"), defaults=dict(a=1.0, b=3.0), name="brusselator")
