/*
 * desolver_oracle.h -- CPU restatement of the Microno95/desolver v5.1.0 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under desolver_b200/ may include, link or
 * call this; it is used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py as the CHECKER for the CUDA path.
 *
 * Parity pin: validated against the unmodified reference (imported through the
 * dispatch-only autoray stand-in in oracle/refshim) by oracle/make_golden.py; the
 * resulting fixtures live in tests/golden/ and tests/test_oracle_golden.py checks
 * this file against them (step counts identical, final states to <=1e-12).
 *
 * Reference lines restated (relative to /root/reference/desolver/):
 *   differential_system.py:951-1002,1015-1018,1070-1071   integrate loop, final-step clamp
 *   integrators/integrator_types.py:162-228               attempt + <=64 retries
 *   integrators/integrator_types.py:267-323               step, dState, error estimate
 *   integrators/components/runge_kutta_methods.py:44-54   stage sums (left-to-right)
 *   integrators/integrator_template.py:46-93              step-size controller
 *   integrators/integrator_types.py:402-417               symplectic drift/kick loop
 *   utilities/interpolation.py:41-61                      cubic Hermite evaluation
 *   backend/autoray_backend.py:8-31                       4*eps / 32*eps thresholds
 */
#ifndef DESOLVER_ORACLE_H
#define DESOLVER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Built-in right-hand sides.  Numbering is shared with include/desolver_b200.h. */
enum {
    ORC_RHS_LINEAR = 0, /* dy = A @ y, A (dim x dim) row-major in shared[]            */
    ORC_RHS_LORENZ = 1, /* traj params: sigma, rho, beta                              */
    ORC_RHS_VDP    = 2, /* traj params: mu ; dy = [v, mu*(1-x*x)*v - x]               */
    ORC_RHS_NBODY  = 3, /* state (2, nb, 3) = (q; v); shared = [G, eps2, m_0..m_nb-1] */
    ORC_RHS_SHO    = 4, /* traj params: k, m ; dy = [y1, -(k/m)*y0]                   */
    ORC_RHS_FORCED = 5, /* dy = [y1, -y0 + t]   (time dependent; reference test fixture)*/
    ORC_RHS_DUFFING = 6 /* traj params: delta, alpha, beta, gamma, omega;
                           dy = [v, gamma*cos(omega*t) - delta*v - alpha*x - beta*x*x*x]   */
};

enum {
    ORC_STATUS_OK       = 0,
    ORC_STATUS_TOL_FAIL = 1, /* FailedToMeetTolerances: 64 retries exhausted */
    ORC_STATUS_RUNNING  = 2  /* max_steps budget hit before tf               */
};

typedef struct {
    int rhs_id;
    int dim;               /* flattened state size of ONE trajectory                 */
    int n_params;          /* number of per-trajectory scalar parameters             */
    const double *shared;  /* shared block (see enum above), may be NULL             */
    int n_shared;
    int drift_count;       /* symplectic only: leading state elements that are drift  */
                           /* variables = (dim0 // 2) * prod(dim[1:]); < 0: dim / 2   */
} orc_problem;

typedef struct {
    int stages;            /* S                                                       */
    const double *inter;   /* S x (S+1) row-major: column 0 = c, columns 1.. = a      */
    const double *final_;  /* final_rows x (S+1): row 0 propagated, row 1 embedded    */
    int final_rows;        /* 1 (fixed step) or 2 (adaptive)                          */
    double order;          /* __order__                                               */
} orc_rk_tableau;

/* Optional per-accepted-step node recorder (dense output): t0,t1,y0,y1,f0,f1. */
typedef struct {
    int64_t capacity;      /* number of steps the arrays can hold                     */
    int64_t count;         /* steps recorded (may exceed capacity => truncated)       */
    double *t0, *t1;       /* [capacity]                                              */
    double *y0, *y1, *f0, *f1; /* [capacity][dim]                                     */
} orc_dense;

/* One trajectory, explicit Runge-Kutta, reference semantics.  y[dim], *t, *dt are
 * in/out.  params[n_params].  Returns an ORC_STATUS_* code. */
int orc_erk_trajectory(const orc_problem *prob, const orc_rk_tableau *tab,
                       const double *params, double *y, double *t, double *dt,
                       double tf, double rtol, double atol, int64_t max_steps,
                       int64_t *accepted, int64_t *attempts, orc_dense *dense);

/* One system, explicit symplectic splitting (fixed dt).  tab3 is S x 3 row-major. */
int orc_symplectic_trajectory(const orc_problem *prob, int stages, const double *tab3,
                              const double *params, double *y, double *t, double *dt,
                              double tf, int64_t max_steps, int64_t *steps);

/* Ensemble drivers over SoA arrays: y[dim][n], t[n], dt[n], params[n_params][n].
 * Trajectories first..first+count-1 are processed (so callers can shard by thread). */
void orc_erk_ensemble(const orc_problem *prob, const orc_rk_tableau *tab,
                      const double *params, double *y, double *t, double *dt,
                      int64_t n, int64_t first, int64_t count,
                      double tf, double rtol, double atol,
                      int32_t *accepted, int32_t *rejected, int32_t *status);

void orc_symplectic_ensemble(const orc_problem *prob, int stages, const double *tab3,
                             const double *params, double *y, double *t, double *dt,
                             int64_t n, int64_t first, int64_t count, double tf,
                             int32_t *steps, int32_t *status);

/* Cubic Hermite evaluation of one recorded step at time tq (interpolation.py:41-61). */
void orc_hermite_eval(int dim, double t0, double t1, const double *p0, const double *p1,
                      const double *m0, const double *m1, double tq, double *out);

/* Lookup + evaluation over a recorded trajectory (differential_system.py:205-227). */
void orc_dense_eval(const orc_dense *dense, int dim, double tq, double *out);

/* Plain RHS evaluation (exposed for tests). */
void orc_rhs(const orc_problem *prob, const double *params, double t, const double *y, double *dy);

/* Test knob: every step-size factor the controller returns is moved by `ulps` units in the last place (0 = off).
 * Used by tests/test_oracle_golden.py to show that the parity tail (a handful of 2^20 chaotic trajectories beyond
 * 1e-10) is the amplification of last-ulp differences, not an arithmetic error.  Not thread-safe: set it between runs. */
void orc_set_factor_perturbation(int ulps);

#ifdef __cplusplus
}
#endif
#endif
