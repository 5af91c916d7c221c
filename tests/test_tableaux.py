"""Tableau data and the method registry against the reference's class data (tests/golden/tableaux.npz,
method_names.json written by oracle/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

from conftest import GOLDEN, load_golden
from desolver_b200 import integrators
from desolver_b200.integrators import tableaux as tb


def test_tableaux_bit_identical_to_reference():
    g = load_golden("tableaux.npz")
    names = sorted({k.split(".")[0] for k in g})
    assert len(names) == 16
    for name in names:
        cls = getattr(integrators, name)
        assert cls.tableau_intermediate.tobytes() == g[name + ".inter"].tobytes(), name
        assert float(cls.__order__) == float(g[name + ".order"])
        if name + ".final" in g:
            assert cls.tableau_final.tobytes() == g[name + ".final"].tobytes(), name
            assert tb.rk_arrays(name)[0].tobytes() == g[name + ".inter"].tobytes()


def test_registry_names_match_reference():
    reg = json.load(open(os.path.join(GOLDEN, "method_names.json")))
    avail = integrators.available_methods(False)
    for cls_name, alts in reg.items():
        assert avail[cls_name].__name__ == cls_name
        for alt in alts:
            assert avail[alt].__name__ == cls_name
    assert integrators.available_methods() == sorted(set(avail.keys()))
    assert integrators.available_methods(False)["RK45"] is integrators.RK45CKSolver
    assert integrators.available_methods(False)["RK87"] is integrators.RK8713MSolver


def test_integrator_flags_like_reference():
    # reference integrators/tests/test_integrators.py:6-21: defaults 32*eps, adaptive/explicit/FSAL flags
    for cls in integrators.explicit_methods():
        a = cls((3,) if not cls.symplectic else (2, 3), "float64")
        assert a.rtol == a.atol == 32 * 4 * np.finfo(np.float64).eps
        assert a.is_explicit and not a.is_implicit
        if not cls.symplectic:
            assert a.is_adaptive == (cls.tableau_final.shape[0] == 2)
            assert a.is_fsal == bool(np.all(cls.tableau_intermediate[-1, 1:] == cls.tableau_final[0, 1:]))
    assert integrators.DOPRI45((3,), "float64").is_fsal and not integrators.RK45CKSolver((3,), "float64").is_fsal
