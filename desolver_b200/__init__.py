"""desolver_b200 -- B200-native ensemble ODE engine behind the desolver Python API.

    import desolver_b200 as de
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    a.method = "RK45"
    a.integrate()

mirrors `import desolver as de` of the reference (/root/reference/desolver/__init__.py:1-7).
Importing the package does not need a GPU; constructing an `OdeSystem` does, and raises if
libdesolver_b200.so is not built (there is no CPU fallback).
"""
from . import backend
from . import exception_types
from . import integrators
from . import rhs_plugins
from . import benchmarks
from . import distributed
from . import events
from . import utilities
from . import differential_system
from .differential_system import DiffRHS, OdeSystem, DenseOutput, StateTuple, rhs_prettifier, solve_ivp
from .integrators import available_methods

__version__ = "0.1.0"
__all__ = ["backend", "exception_types", "integrators", "rhs_plugins", "benchmarks", "distributed", "events", "DiffRHS", "OdeSystem",
           "DenseOutput", "StateTuple", "rhs_prettifier", "solve_ivp", "available_methods"]
