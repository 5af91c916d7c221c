"""Method registry of the drop-in (reference: desolver/integrators/__init__.py:10-72).

Every explicit scheme of the reference is available under the same class name and the same
`__alt_names__`; the classes are generated from the tableau data.  Implicit schemes are out of
scope of the ensemble engine and are not registered.
"""
from . import tableaux as _tb
from ._tableau_data import EXPLICIT_RK as _RK, SYMPLECTIC as _SY
from .integrator_types import (ExplicitSymplecticIntegrator, IntegratorTemplate, RungeKuttaIntegrator,
                               TableauIntegrator)
from .richardson import RichardsonIntegratorTemplate, generate_richardson_integrator

__explicit_integration_methods__ = []
__available_methods = {}


def _make_classes():
    order_of_definition = ["RK1412Solver", "RK108Solver", "RK8713MSolver", "RK45CKSolver", "RK5Solver", "RK4Solver",
                           "MidpointSolver", "HeunsSolver", "RalstonsSolver", "EulerSolver", "EulerTrapSolver",
                           "HeunEulerSolver", "SymplecticEulerSolver", "BABs9o7HSolver", "ABAs5o6HSolver", "DOPRI45"]
    for name in order_of_definition:
        if name in _RK:
            inter, final, order = _tb.rk_arrays(name)
            body = dict(tableau_intermediate=inter, tableau_final=final, __order__=order,
                        __alt_names__=_tb.alt_names(name), __doc__="Explicit Runge-Kutta scheme %s" % name)
            cls = type(name, (RungeKuttaIntegrator,), body)
        else:
            tab, order = _tb.symplectic_array(name)
            body = dict(tableau_intermediate=tab, __order__=order, __alt_names__=_tb.alt_names(name),
                        __doc__="Explicit symplectic splitting scheme %s" % name)
            cls = type(name, (ExplicitSymplecticIntegrator,), body)
        cls.__module__ = __name__
        globals()[name] = cls
        __explicit_integration_methods__.append(cls)
        __available_methods[name] = cls
        for alt in cls.__alt_names__:
            __available_methods[alt] = cls


_make_classes()


def available_methods(names=True):
    if names:
        return sorted(set(__available_methods.keys()))
    return __available_methods


def explicit_methods():
    return __explicit_integration_methods__


def implicit_methods():
    """Implicit Runge-Kutta schemes (Gauss-Legendre, Lobatto, Radau, DIRK, Backward Euler ...) are outside the hot path
    this engine replaces (batched Newton + dense LU per stage: a different algorithmic class);
    none is registered and `OdeSystem.set_method` names the reference class to use instead."""
    return []


# names of the reference's implicit / wrapper schemes, so that asking for one gets a targeted error instead of
# "method does not exist"
REFERENCE_ONLY_METHODS = ("GaussLegendre4", "GaussLegendre6", "BackwardEuler", "ImplicitMidpoint", "LobattoIIIA2",
                          "LobattoIIIA4", "LobattoIIIB2", "LobattoIIIB4", "LobattoIIIC2", "LobattoIIIC4", "CrankNicolson",
                          "DIRK3LStable", "RadauIA3", "RadauIA5", "RadauIIA3", "RadauIIA5", "RadauIIA19")


def register_with_reference(reference_integrators=None):
    """Seam 1 of SURVEY.md section 8b on the reference's side: the unmodified reference `OdeSystem.set_method` accepts a
    class only if `issubclass(cls, desolver.integrators.IntegratorTemplate)` (differential_system.py:858-866).  The
    reference's template is an `abc.ABC`, so registering this package's template as a virtual subclass makes every class
    here acceptable without touching either code base:

        import desolver, desolver_b200
        desolver_b200.integrators.register_with_reference(desolver.integrators)
        a = desolver.OdeSystem(rhs, y0_cuda, ...); a.set_method(desolver_b200.integrators.RK45CKSolver)
    """
    if reference_integrators is None:
        import desolver.integrators as reference_integrators
    reference_integrators.IntegratorTemplate.register(IntegratorTemplate)
    return reference_integrators


__all__ = ["IntegratorTemplate", "TableauIntegrator", "RungeKuttaIntegrator", "ExplicitSymplecticIntegrator",
           "available_methods", "explicit_methods", "implicit_methods", "register_with_reference", "RichardsonIntegratorTemplate",
           "generate_richardson_integrator"] + \
          [c.__name__ for c in __explicit_integration_methods__]
