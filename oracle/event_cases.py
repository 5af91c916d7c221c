"""Non-affine event functions shared by oracle/make_golden_events2.py (run through the unmodified reference) and
tests/test_gpu_events_general.py (run through the event-search kernels).  TEST INFRASTRUCTURE.

Every function follows the reference's protocol `event(t, y[, dy], **constants)` with the attributes `is_terminal`,
`direction`, `requires_dstate` (/root/reference/desolver/differential_system.py:29-57) and is written so that it works
on numpy arrays of shape (dim,) (the reference, one system at a time) and on torch tensors of shape (dim, N) (the
ensemble engine) alike."""
import math


def _ev(direction=0, is_terminal=False, requires_dstate=False):
    def deco(fn):
        fn.direction, fn.is_terminal, fn.requires_dstate = direction, is_terminal, requires_dstate
        return fn
    return deco


@_ev(direction=0)
def lorenz_sphere(t, y, **constants):
    """crossings of the sphere of radius 18 around (0, 0, 25)"""
    return y[0] * y[0] + y[1] * y[1] + (y[2] - 25.0) * (y[2] - 25.0) - 324.0


@_ev(direction=-1)
def lorenz_product(t, y, **constants):
    """x y falls through 40"""
    return y[0] * y[1] - 40.0


@_ev(direction=1, is_terminal=True)
def lorenz_sphere_terminal(t, y, **constants):
    """stop when the trajectory LEAVES the sphere of radius 22 around (0, 0, 25)"""
    return y[0] * y[0] + y[1] * y[1] + (y[2] - 25.0) * (y[2] - 25.0) - 484.0


@_ev(direction=0)
def vdp_energy_mu(t, y, mu=1.0, **constants):
    """x^2 + v^2 crosses a level that depends on the per-trajectory constant mu"""
    return y[0] * y[0] + y[1] * y[1] - (3.0 + 0.2 * mu)


@_ev(direction=-1, is_terminal=True, requires_dstate=True)
def radius_maximum(t, y, dy, **constants):
    """d/dt (x^2 + v^2) / 2 = x x' + v v' falls through zero: a maximum of the radius (sees the interpolant's derivative)"""
    return y[0] * dy[0] + y[1] * dy[1]


@_ev(direction=0)
def time_cubic(t, y, **constants):
    """the reference's own multiple-roots test (tests/test_event_detection.py:138-146): three roots in t"""
    return (t - math.pi / 8) * (t - math.pi / 16) * (t - math.pi / 32) + 0.0 * y[0]


CASES = {
    # name: (rhs name, events, tf, rtol = atol, method)
    "lorenz_nonaffine": ("lorenz", [lorenz_sphere, lorenz_product], 2.0, 1e-9, "RK45"),
    "lorenz_nonaffine_terminal": ("lorenz", [lorenz_sphere_terminal, lorenz_product], 2.0, 1e-9, "RK45"),
    "vdp_energy_mu": ("van_der_pol", [vdp_energy_mu], 10.0, 1e-9, "RK45"),
    "vdp_radius_maximum": ("van_der_pol", [radius_maximum, vdp_energy_mu], 10.0, 1e-9, "RK45"),
    "vdp_energy_mu_rk87": ("van_der_pol", [vdp_energy_mu], 10.0, 1e-9, "RK87"),
}
