"""Exception types of the drop-in.  The NAMES are part of the API surface: user code catches
`desolver.exception_types.FailedIntegration` and inspects `__cause__` (reference
/root/reference/desolver/exception_types/exception_types.py:1-13, raised at differential_system.py:1089-1093 and
integrators/integrator_types.py:220-224).  In the ensemble engine the per-trajectory status word written by the
kernels is mapped onto them after the run (error policy)."""

__all__ = ["RecursionError", "FailedIntegration", "FailedToMeetTolerances"]


class FailedToMeetTolerances(Exception):
    """At least one trajectory exhausted its 64 step retries (status DSB_TRAJ_TOL_FAIL)."""


class FailedIntegration(Exception):
    """Raised by `OdeSystem.integrate`; the underlying reason is attached as `__cause__`."""


class RecursionError(Exception):  # noqa: A001  (shadows the builtin on purpose, like the reference module)
    """Kept for source compatibility; the ensemble engine never recurses."""
