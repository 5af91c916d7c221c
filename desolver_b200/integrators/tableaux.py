"""Dense Butcher / splitting tableaux in the reference's array layout.

`tableau_intermediate` is S x (S+1) with column 0 = c and columns 1.. = a; `tableau_final`
is (1|2) x (S+1) with row 0 the propagated weights and row 1 the embedded (error) weights
-- the layout of /root/reference/desolver/integrators/explicit_integration_schemes.py:24-449,
so the integrator classes expose the same class attributes as the reference's.  The numbers
come from `_tableau_data` (sparse rational strings) and are bit-identical to the reference's
doubles (pinned by tests/golden/tableaux.npz).
"""
import numpy as np

from ._tableau_data import EXPLICIT_RK, SYMPLECTIC


def _value(txt):
    """'num/den' -> float(num)/float(den); plain literal -> float."""
    txt = str(txt)
    if "/" in txt:
        num, den = txt.split("/")
        return float(num) / float(den)
    return float(txt)


def rk_arrays(class_name):
    """(tableau_intermediate, tableau_final, order) as float64 arrays for an explicit RK class."""
    spec = EXPLICIT_RK[class_name]
    S = spec["stages"]
    inter = np.zeros((S, S + 1), dtype=np.float64)
    for s in range(S):
        inter[s, 0] = _value(spec["c"][s])
        for j, v in spec["a"].get(s, {}).items():
            inter[s, 1 + j] = _value(v)
    rows = 1 if spec["b_embedded"] is None else 2
    final = np.zeros((rows, S + 1), dtype=np.float64)
    for r in range(rows):
        final[r, 0] = _value(spec["b_col0"][r])
    for j, v in spec["b"].items():
        final[0, 1 + j] = _value(v)
    if rows == 2:
        for j, v in spec["b_embedded"].items():
            final[1, 1 + j] = _value(v)
    return inter, final, float(spec["order"])


def symplectic_array(class_name):
    """(S x 3 tableau, order): column 1 = drift weight (also the time advance), column 2 = kick."""
    spec = SYMPLECTIC[class_name]
    tab = np.array([[_value(a), _value(b), _value(c)]
                    for a, b, c in zip(spec["col0"], spec["drift"], spec["kick"])], dtype=np.float64)
    return tab, float(spec["order"])


def alt_names(class_name):
    spec = EXPLICIT_RK.get(class_name) or SYMPLECTIC[class_name]
    return tuple(spec["alt_names"])
