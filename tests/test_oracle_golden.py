"""Pins the CPU oracle (oracle/desolver_oracle.c) to the unmodified reference: every fixture in
tests/golden/ was produced by the real desolver v5.1.0 (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_inf
from desolver_b200.integrators import tableaux as tb
from oracle import oracle as orc

LORENZ_P = [10.0, 28.0, 8.0 / 3.0]


def _erk(name, cls, rhs_id, params=None, shared=None, **over):
    g = load_golden(name)
    inter, final, order = tb.rk_arrays(cls)
    n = g["y0"].shape[1]
    tf = over.get("tf", g["tf"])
    if np.ndim(tf) == 0 and np.ndim(g["dt0"]) == 0 and np.ndim(g["rtol"]) == 0:
        r = orc.erk_ensemble(rhs_id, g["y0"], 0.0, float(tf), float(g["dt0"]), float(g["rtol"]), float(g["atol"]),
                             inter, final, order, params=params(g) if callable(params) else params,
                             shared=shared(g) if callable(shared) else shared, threads=4)
    else:
        cols = []
        for i in range(n):
            cols.append(orc.erk_ensemble(rhs_id, g["y0"][:, i], 0.0, float(np.broadcast_to(tf, (n,))[i]),
                                         float(np.broadcast_to(g["dt0"], (n,))[i]),
                                         float(np.broadcast_to(g["rtol"], (n,))[i]),
                                         float(np.broadcast_to(g["atol"], (n,))[i]), inter, final, order,
                                         params=params))
        r = {k: (np.stack([c[k] for c in cols], axis=-1) if k == "y" else np.concatenate([c[k] for c in cols]))
             for k in cols[0]}
    return g, r


def _check(g, r, state_tol=1e-12):
    ok = g["status"] == 0
    assert np.array_equal(r["status"][ok], g["status"][ok])
    assert np.array_equal(r["accepted"][ok], g["accepted"][ok]), "accepted step counts differ"
    assert np.array_equal(r["accepted"][ok] + r["rejected"][ok], g["attempts"][ok]), "attempt counts differ"
    assert np.max(rel_inf(r["y"][:, ok], g["y_final"][:, ok])) <= state_tol
    assert np.max(np.abs(r["t"][ok] - g["t_final"][ok])) <= 1e-13
    # failed reference runs must fail in the oracle too
    assert np.all(r["status"][~ok] == orc.STATUS_TOL_FAIL)
    return float(np.mean(np.all(r["y"][:, ok] == g["y_final"][:, ok], axis=0)))


def test_lorenz_rk45_matches_reference():
    g, r = _erk("lorenz_rk45.npz", "RK45CKSolver", orc.RHS_LORENZ, params=LORENZ_P)
    bitwise = _check(g, r)
    assert bitwise > 0.9          # the rest differ only by numpy-SIMD vs glibc atan last-ulp
    # dt is an ill-conditioned function of the state (err is a 1e-9-sized difference of sums), so
    # a last-ulp atan difference moves it by ~1e-9 relative; bitwise runs have bitwise dt
    same = np.all(r["y"] == g["y_final"], axis=0)
    assert np.array_equal(r["dt"][same], g["dt_final"][same])
    assert np.max(np.abs(r["dt"] - g["dt_final"]) / np.abs(g["dt_final"])) < 1e-6
    print("bitwise-identical final states: %.1f %%" % (100 * bitwise))


def test_lorenz_rk87_matches_reference():
    g, r = _erk("lorenz_rk87.npz", "RK8713MSolver", orc.RHS_LORENZ, params=LORENZ_P)
    _check(g, r)


def test_vdp_rk45_matches_reference():
    g, r = _erk("vdp_rk45.npz", "RK45CKSolver", orc.RHS_VDP, params=lambda g: g["mu"][None, :])
    _check(g, r, state_tol=1e-11)


@pytest.mark.parametrize("name,cls", [("lorenz_dopri45.npz", "DOPRI45"), ("lorenz_heun_euler.npz", "HeunEulerSolver"),
                                      ("lorenz_rk4.npz", "RK4Solver"), ("lorenz_rk108.npz", "RK108Solver"),
                                      ("lorenz_rk1412.npz", "RK1412Solver")])
def test_other_tableaux_match_reference(name, cls):
    g, r = _erk(name, cls, orc.RHS_LORENZ, params=LORENZ_P)
    _check(g, r, state_tol=5e-12)


def test_duffing_rk45_matches_reference():
    """Time-dependent right-hand side with per-trajectory forcing (cos through libm vs numpy)."""
    g = load_golden("duffing_rk45.npz")
    n = g["y0"].shape[1]
    prm = np.stack([np.full(n, 0.2), np.full(n, -1.0), np.full(n, 1.0), g["gamma"], g["omega"]])
    r = orc.erk_ensemble(orc.RHS_DUFFING, g["y0"], 0.0, 10.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"), params=prm,
                         threads=4)
    assert np.array_equal(r["accepted"], g["accepted"]) and np.array_equal(r["accepted"] + r["rejected"], g["attempts"])
    assert rel_inf(r["y"], g["y_final"]).max() <= 1e-12


def test_edge_cases_match_reference():
    # oversized dt0 (double rejections -> 3-term controller), dt0 > span clamp, tiny spans
    g, r = _erk("edge_rk45.npz", "RK45CKSolver", orc.RHS_LORENZ, params=LORENZ_P)
    assert (g["attempts"] - g["accepted"]).max() >= 2
    _check(g, r)


def test_linear_rk87_matches_reference():
    g, r = _erk("linear_rk87.npz", "RK8713MSolver", orc.RHS_LINEAR, shared=lambda g: g["A"])
    # early RK8(7) steps at dt0=1e-2 have error estimates of pure rounding noise (~1e-13 against
    # tol 1e-9), so dt -- and through the truncation error the state -- depends on the matvec
    # summation order at the 1e-4 / 1e-12 level; OpenBLAS dgemv order is not reproduced.
    _check(g, r, state_tol=1e-11)


@pytest.mark.parametrize("method,cls", [("BABS9O7H", "BABs9o7HSolver"), ("ABAS5O6H", "ABAs5o6HSolver")])
def test_symplectic_nbody_matches_reference(method, cls):
    g = load_golden("nbody_symplectic.npz")
    tab3, _ = tb.symplectic_array(cls)
    st = g["state0"]                                   # (2, B, nb, 3)
    B = st.shape[1]
    y0 = np.ascontiguousarray(np.moveaxis(st, 1, -1).reshape(-1, B))   # (2*nb*3, B)
    shared = np.concatenate([[float(g["G"]), float(g["eps2"])], g["m"]])
    r = orc.symplectic_ensemble(orc.RHS_NBODY, y0, 0.0, float(g["tf"]), float(g["dt"]), tab3, shared=shared)
    ref = np.moveaxis(g[method + ".y"], 1, -1).reshape(-1, B)
    assert np.all(r["steps"] == len(g[method + ".t"]) - 1)
    assert np.max(rel_inf(r["y"], ref)) < 1e-13
    assert np.allclose(r["t"], g[method + ".t"][-1], rtol=0, atol=1e-15)


def test_known_answers():
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    hx = lambda v: np.array([float.fromhex(s) for s in v])
    ck = tb.rk_arrays("RK45CKSolver")
    dp = tb.rk_arrays("RK8713MSolver")
    two_pi = float.fromhex(ka["KA0"]["t"][0])
    for key, tab in (("KA0", ck), ("KA3", dp)):
        r = orc.erk_trajectory(orc.RHS_SHO, [1.0, 0.0], 0.0, two_pi, 0.01, 1e-9, 1e-9, *tab, params=[1.0, 1.0])
        assert r["accepted"] == ka[key]["accepted"]
        # nfev = 1 (constructor shape check) + 1 (first initial_rhs) + (S+1) per attempt
        assert 2 + (tab[0].shape[0] + 1) * r["attempts"] == ka[key]["nfev"]
        assert np.max(np.abs(r["y"] - hx(ka[key]["y"]))) < 1e-15
        assert r["t"] == two_pi
    r = orc.erk_trajectory(orc.RHS_LORENZ, [1.0, 1.0, 20.0], 0.0, 2.0, 1e-2, 1e-9, 1e-9, *ck, params=LORENZ_P)
    assert r["accepted"] == ka["KA1"]["accepted"] == 256 and 2 + 7 * r["attempts"] == ka["KA1"]["nfev"] == 1801
    assert np.max(np.abs(r["y"] - hx(ka["KA1"]["y"])) / np.abs(hx(ka["KA1"]["y"]))) < 1e-13
    for key, mu in (("KA2a", 0.1), ("KA2b", 10.0)):
        r = orc.erk_trajectory(orc.RHS_VDP, [2.0, 0.0], 0.0, 10.0, 1e-2, 1e-9, 1e-9, *ck, params=[mu])
        assert r["accepted"] == ka[key]["accepted"] and 2 + 7 * r["attempts"] == ka[key]["nfev"]
        assert np.max(np.abs(r["y"] - hx(ka[key]["y"]))) < 1e-12
    for key, cls in (("KA4", "BABs9o7HSolver"), ("KA5", "ABAs5o6HSolver")):
        tab3, _ = tb.symplectic_array(cls)
        r = orc.symplectic_ensemble(orc.RHS_SHO, np.array([1.0, 0.0]), 0.0, two_pi, 0.05, tab3, params=[1.0, 1.0])
        assert r["steps"][0] == ka[key]["steps"] == 126
        assert np.array_equal(r["y"], hx(ka[key]["y"]))              # fixed step: bit-exact
        assert r["t"][0] == hx(ka[key]["t"])[0]


def test_dense_output_matches_reference():
    g = load_golden("lorenz_dense.npz")
    ck = tb.rk_arrays("RK45CKSolver")
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for i in range(g["y0"].shape[1]):
        r = orc.erk_trajectory(orc.RHS_LORENZ, g["y0"][:, i], 0.0, 2.0, 1e-2, 1e-9, 1e-9, *ck, params=LORENZ_P,
                               dense_capacity=512)
        assert r["dense"].count == g["n_interp"][i] == g["accepted"][i]
        got = np.stack([r["dense"](tq) for tq in g["tq"]])
        assert np.max(np.abs(got - g["dense"][:, :, i])) < 1e-11
    r = orc.erk_trajectory(orc.RHS_SHO, [1.0, 0.0], 0.0, float.fromhex(ka["KA0"]["t"][0]), 0.01, 1e-9, 1e-9, *ck,
                           params=[1.0, 1.0], dense_capacity=128)
    assert r["dense"].count == ka["KA6"]["n_interp"] == 84
    for tq, hexes in ka["KA6"]["sol_at"]:
        want = np.array([float.fromhex(s) for s in hexes])
        assert np.max(np.abs(r["dense"](tq) - want)) < 1e-15


def test_parity_tail_is_chaos_not_arithmetic():
    """The full 2^20-trajectory comparison (profiles/round1_parity_distribution_full_2pow20.txt) has every step count
    identical, a median state error of 2.7e-15 and p99 of 1.3e-13 -- and 5 trajectories between 1e-10 and 8e-10.
    Moving the ORACLE's OWN step-size factor by one unit in the last place (the difference between two correct
    libm's) reproduces exactly that picture: the same trajectories (indices recorded by the GPU
    run) move by 5e-11 ... 8e-10, the bulk by a median of ~3e-15 with p99 ~1e-13, and no step count changes.  The tail is
    therefore the Lorenz flow amplifying last-ulp noise, which no implementation can remove."""
    from desolver_b200 import benchmarks as bm
    cfg = bm.LORENZ
    y0 = bm.lorenz_ensemble(1 << 20)
    worst = [635367, 968982, 199215, 527174]                 # GPU-vs-oracle errors 1.9e-10 ... 7.8e-10 (fast and STRICT)
    idx = np.array(worst + list(range(8192)))
    tab = tb.rk_arrays("RK45CKSolver")

    def run(ulps):
        orc.set_factor_perturbation(ulps)
        try:
            return orc.erk_ensemble(orc.RHS_LORENZ, np.ascontiguousarray(y0[:, idx]), 0.0, cfg["tf"], cfg["dt0"], 1e-9, 1e-9, *tab,
                                    params=[10.0, 28.0, 8.0 / 3.0], threads=os.cpu_count())
        finally:
            orc.set_factor_perturbation(0)
    base = run(0)
    again = run(0)
    assert np.array_equal(base["y"], again["y"])              # the knob is off again, the oracle is deterministic
    tail_moves = []
    for ulps in (1, -1):
        r = run(ulps)
        assert np.array_equal(r["accepted"], base["accepted"]) and np.array_equal(r["rejected"], base["rejected"])
        err = rel_inf(r["y"], base["y"])
        bulk = err[len(worst):]
        assert np.median(bulk) < 1e-14 and np.percentile(bulk, 99) < 1e-12      # GPU vs oracle: 2.7e-15 / 1.3e-13
        tail_moves.append(err[:len(worst)])
    tail = np.max(np.stack(tail_moves), axis=0)               # per trajectory: the larger of the two responses
    assert np.all(tail > 5e-11) and np.all(tail < 5e-9), tail  # same magnitude as what the GPU run shows for them
    assert tail.min() > 100 * np.percentile(bulk, 99)         # ... and far outside the bulk
