"""Seam 1 of SURVEY.md section 8b, checked against the UNMODIFIED reference in the build container (no GPU): the
reference's own `OdeSystem.set_method` accepts this package's integrator classes (constructor signature, class and
instance attributes), and the protocol-level controller `update_timestep` reproduces the reference's arithmetic.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "desolver")), reason="reference sources not present")


@pytest.fixture(scope="module")
def ref():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF]
    import desolver
    import desolver_b200
    desolver_b200.integrators.register_with_reference(desolver.integrators)
    return desolver


def test_reference_odesystem_accepts_the_drop_in_classes(ref):
    import desolver_b200 as b2

    def rhs(t, y, **kwargs):
        return -y
    a = ref.OdeSystem(rhs, y0=np.array([1.0, 0.0]), t=(0, 1), dt=0.01, rtol=1e-7, atol=1e-8)
    for cls in b2.integrators.explicit_methods():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a.set_method(cls)                                   # differential_system.py:839-869
        integ = a.integrator
        assert type(integ) is cls and a.method is cls
        refcls = getattr(ref.integrators, cls.__name__)
        assert integ.order == refcls.__order__ and integ.symplectic == refcls.symplectic
        assert integ.rtol == 1e-7 and integ.atol == 1e-8 and integ.dim == (2,)
        r = refcls((2,), dtype=np.float64, rtol=1e-7, atol=1e-8)
        assert integ.is_adaptive == r.is_adaptive and integ.is_implicit == r.is_implicit and integ.is_fsal == r.is_fsal
        assert integ.stages == r.stages
        if cls.symplectic:
            assert np.array_equal(np.asarray(integ.staggered_mask), np.asarray(r.staggered_mask))
        else:
            assert set(r.solver_dict_keep_keys) <= set(integ.solver_dict.keys()) | {"num_step_retries"}
        for attr in ("dense_output", "update_timestep", "adaptation_fn", "dState", "dTime", "solver_dict"):
            assert hasattr(integ, attr)
    # default tolerances: 32 * D.epsilon(dtype) (integrator_types.py:37-38)
    d = b2.integrators.RK45CKSolver((3,), np.float64)
    r = ref.integrators.RK45CKSolver((3,), dtype=np.float64)
    assert d.rtol == float(r.rtol) and d.atol == float(r.atol)


def test_update_timestep_matches_the_reference_controller(ref):
    """Same solver_dict in, same (timestep, redo) out -- first attempt, two-term and three-term retry branches
    (integrator_template.py:46-93)."""
    import torch
    import desolver_b200 as b2
    rng = np.random.default_rng(5)
    mine = b2.integrators.RK45CKSolver((3,), "float64", rtol=1e-9, atol=1e-9)
    theirs = ref.integrators.RK45CKSolver((3,), dtype=np.float64, rtol=1e-9, atol=1e-9)
    for scale in (1e-10, 1e-8, 1e-6):                          # accepted, borderline, rejected
        y = rng.normal(size=3)
        dstate = rng.normal(size=3) * 1e-2
        for trial in range(4):
            diff = rng.normal(size=3) * scale
            h = 0.01 / (trial + 1)
            theirs.solver_dict.update(initial_state=y.copy(), diff=diff.copy(), timestep=np.float64(h), dState=dstate.copy(),
                                      atol=1e-9, rtol=1e-9)
            mine._seed_solver_dict()
            mine.solver_dict.update(initial_state=torch.as_tensor(y), diff=torch.as_tensor(diff), timestep=torch.as_tensor(h, dtype=torch.float64),
                                    dState=torch.as_tensor(dstate), atol=1e-9, rtol=1e-9)
            t_ref, redo_ref = theirs.update_timestep()
            t_mine, redo_mine = mine.update_timestep()
            assert redo_ref == redo_mine
            assert abs(float(t_mine) - float(t_ref)) <= 4e-16 * abs(float(t_ref))
        for sd in (mine.solver_dict, theirs.solver_dict):       # a new step wipes the controller history (:163)
            for k in ("system_scaling", "epsilon_last", "epsilon_last_last"):
                sd.pop(k, None)


def test_custom_adaptation_fn_protocol(ref):
    import desolver_b200 as b2
    integ = b2.integrators.RK45CKSolver((2,), "float64", rtol=1e-6, atol=1e-6)
    assert not integ.has_custom_adaptation and integ.adaptation_fn == integ.update_timestep

    def halve(self):
        return 0.5 * self.solver_dict["timestep"], False
    integ.adaptation_fn = halve
    assert integ.has_custom_adaptation and integ.adaptation_fn is halve
    with pytest.raises(ValueError):
        integ.adaptation_fn = lambda self: 1.0                  # not a (timestep, bool) tuple


def test_reference_only_methods_raise_a_targeted_error():
    import desolver_b200 as b2
    assert b2.integrators.implicit_methods() == []
    assert "RadauIIA5" in b2.integrators.REFERENCE_ONLY_METHODS


def test_installed_reference_is_the_unmodified_tree():
    """baseline/_ref (the copy that travels to the GPU box for bench.py's live CPU-A / CPU-B timing and
    `parity.reference_live`) is byte for byte /root/reference/desolver."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import install_reference
    tree = install_reference.install()
    assert tree is not None and os.path.isdir(os.path.join(tree, "desolver"))
    assert install_reference._same_tree(os.path.join(REF, "desolver"), os.path.join(tree, "desolver"))


def test_utilities_lookups_equal_the_references(ref):
    """`desolver_b200.utilities.search_bisection(_vec)` against `desolver.utilities` on sorted arrays, including the edges
    and exact hits."""
    import desolver_b200 as b2
    rng = np.random.default_rng(0)
    for _ in range(20):
        arr = np.sort(rng.uniform(-3, 3, int(rng.integers(2, 40))))
        vals = np.concatenate([rng.uniform(-4, 4, 30), arr[:3], [arr[0], arr[-1]]])
        for v in vals:
            assert b2.utilities.search_bisection(arr, v) == ref.utilities.search_bisection(arr, v)
        assert np.array_equal(b2.utilities.search_bisection_vec(arr, vals), ref.utilities.search_bisection_vec(arr, vals))
    assert hasattr(b2, "utilities") and hasattr(b2, "differential_system") and b2.utilities.CubicHermiteInterp is not None
