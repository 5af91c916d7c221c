/*
 * desolver_oracle.c -- CPU restatement of the Microno95/desolver v5.1.0 hot path.
 * TEST INFRASTRUCTURE ONLY (see desolver_oracle.h for the rules and the pin).
 *
 * Build with contraction OFF (oracle/Makefile uses -ffp-contract=off): every +,* below
 * rounds separately, as numpy's separate ufunc calls do; the one place the reference
 * arithmetic fuses (OpenBLAS ddot inside linalg.norm) is written with explicit fma().
 */
#include "desolver_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS4   (4.0 * DBL_EPSILON)   /* D.epsilon      backend/autoray_backend.py:8-18  */
#define ORC_EPS32  (32.0 * DBL_EPSILON)  /* D.tol_epsilon  backend/autoray_backend.py:21-31 */
#define ORC_RETRIES 64                   /* num_step_retries integrator_types.py:148        */

/* np.maximum: NaN-propagating */
static double np_maximum(double a, double b)
{
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    return a > b ? a : b;
}

/* ------------------------------------------------------------------ RHS ---- */

static void rhs_linear(const orc_problem *p, const double *y, double *dy)
{
    int n = p->dim;
    const double *A = p->shared;
    for (int i = 0; i < n; ++i) {
        double s = A[(size_t)i * n] * y[0];
        for (int j = 1; j < n; ++j) s = s + A[(size_t)i * n + j] * y[j];
        dy[i] = s;
    }
}

static void rhs_lorenz(const double *prm, const double *s, double *dy)
{
    double sigma = prm[0], rho = prm[1], beta = prm[2];
    double x = s[0], y = s[1], z = s[2];
    dy[0] = sigma * (y - x);
    dy[1] = x * (rho - z) - y;
    dy[2] = x * y - beta * z;
}

static void rhs_vdp(const double *prm, const double *s, double *dy)
{
    double mu = prm[0], x = s[0], v = s[1];
    dy[0] = v;
    dy[1] = mu * (1.0 - x * x) * v - x;
}

static void rhs_sho(const double *prm, const double *s, double *dy)
{
    double w = -(prm[0] / prm[1]);
    dy[0] = s[1];
    dy[1] = w * s[0];
}

/* state (2, nb, 3): q = y[0 .. 3nb), v = y[3nb .. 6nb) ; shared = [G, eps2, m...] */
static void rhs_nbody(const orc_problem *p, const double *y, double *dy)
{
    int nb = p->dim / 6;
    double G = p->shared[0], eps2 = p->shared[1];
    const double *m = p->shared + 2;
    const double *q = y, *v = y + 3 * nb;
    for (int i = 0; i < 3 * nb; ++i) dy[i] = v[i];
    for (int i = 0; i < nb; ++i) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < nb; ++j) {
            double dx = q[3 * j] - q[3 * i], dyy = q[3 * j + 1] - q[3 * i + 1], dz = q[3 * j + 2] - q[3 * i + 2];
            double r2 = ((dx * dx + dyy * dyy) + dz * dz) + eps2;
            double w = (i == j) ? 0.0 : m[j] * pow(r2, -1.5);
            if (j == 0) { ax = dx * w; ay = dyy * w; az = dz * w; }
            else        { ax = ax + dx * w; ay = ay + dyy * w; az = az + dz * w; }
        }
        dy[3 * nb + 3 * i]     = G * ax;
        dy[3 * nb + 3 * i + 1] = G * ay;
        dy[3 * nb + 3 * i + 2] = G * az;
    }
}

void orc_rhs(const orc_problem *p, const double *prm, double t, const double *y, double *dy)
{
    switch (p->rhs_id) {
    case ORC_RHS_FORCED: dy[0] = y[1]; dy[1] = -y[0] + t; break;
    case ORC_RHS_DUFFING:   /* desolver_b200/rhs_plugins.py: duffing, left to right */
        dy[0] = y[1];
        dy[1] = ((prm[3] * cos(prm[4] * t) - prm[0] * y[1]) - prm[1] * y[0]) - ((prm[2] * y[0]) * y[0]) * y[0];
        break;
    case ORC_RHS_LINEAR: rhs_linear(p, y, dy); break;
    case ORC_RHS_LORENZ: rhs_lorenz(prm, y, dy); break;
    case ORC_RHS_VDP:    rhs_vdp(prm, y, dy); break;
    case ORC_RHS_NBODY:  rhs_nbody(p, y, dy); break;
    case ORC_RHS_SHO:    rhs_sho(prm, y, dy); break;
    default: for (int i = 0; i < p->dim; ++i) dy[i] = NAN;
    }
}

/* -------------------------------------------------- numpy reduce orders ---- */

/* numpy add.reduce over a CONTIGUOUS inner axis of length n (integrator_types.py:307,321):
 * n < 8 left-to-right; otherwise 8 running accumulators over blocks of 8, combined as
 * ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n%8 tail added in order.  (n <= 128 here.) */
static double np_sum_contig(const double *term, int n)
{
    if (n < 8) {
        double s = term[0];
        for (int i = 1; i < n; ++i) s = s + term[i];
        return s;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = term[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = r[j] + term[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + term[i];
    return res;
}

/* linalg.norm of a 1-D float64 vector = sqrt(ddot(x,x)).  OpenBLAS' small-n path is an
 * FMA chain left to right (integrator_template.py:63-64).  For n >= 32 the Haswell
 * micro-kernel keeps 16 partial sums; its exact combine order is not reproduced bit for
 * bit (parity there is tolerance-based; none of the golden configs with dim >= 32 needs it). */
static double l2_norm(const double *v, int n)
{
    if (n < 32) {
        double s = v[0] * v[0];
        for (int i = 1; i < n; ++i) s = fma(v[i], v[i], s);
        return sqrt(s);
    }
    double acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    int n1 = n & -32, i = 0;
    for (; i < n1; i += 16)
        for (int k = 0; k < 16; ++k) acc[k] = fma(v[i + k], v[i + k], acc[k]);
    double lane[4];
    for (int k = 0; k < 4; ++k) lane[k] = (acc[k] + acc[4 + k]) + (acc[8 + k] + acc[12 + k]);
    double s = (lane[0] + lane[1]) + (lane[2] + lane[3]);
    for (; i < n; ++i) s = fma(v[i], v[i], s);
    return sqrt(s);
}

/* ------------------------------------------------- explicit Runge-Kutta ---- */

typedef struct {
    int dim, S;
    double *K;       /* [S][dim] stage derivatives (persist across attempts, like stage_values) */
    double *ystage, *dS, *diff, *scale, *tmp, *terms;
    double *f_init, *f_final;
    int have_final;
} erk_work;

static int erk_work_alloc(erk_work *w, int dim, int S)
{
    w->dim = dim; w->S = S;
    size_t nd = (size_t)dim;
    w->K = calloc((size_t)S * nd, sizeof(double));
    w->ystage = malloc(nd * sizeof(double));
    w->dS = malloc(nd * sizeof(double));
    w->diff = malloc(nd * sizeof(double));
    w->scale = malloc(nd * sizeof(double));
    w->tmp = malloc(nd * sizeof(double));
    w->terms = malloc((size_t)(S > 8 ? S : 8) * sizeof(double));
    w->f_init = malloc(nd * sizeof(double));
    w->f_final = malloc(nd * sizeof(double));
    w->have_final = 0;
    return w->K && w->ystage && w->dS && w->diff && w->scale && w->tmp && w->terms && w->f_init && w->f_final;
}

static void erk_work_free(erk_work *w)
{
    free(w->K); free(w->ystage); free(w->dS); free(w->diff); free(w->scale);
    free(w->tmp); free(w->terms); free(w->f_init); free(w->f_final);
}

/* One attempt: compute_step (runge_kutta_methods.py:44-54) + step (integrator_types.py:302-308)
 * + get_error_estimate (:319-323).  Fills w->dS, w->diff (= h * err) and w->f_final. */
static void erk_attempt(const orc_problem *prob, const orc_rk_tableau *tab, const double *prm,
                        erk_work *w, double t, const double *y, double h, int fsal)
{
    const int S = tab->stages, dim = w->dim, W = S + 1;
    const double *A = tab->inter, *B = tab->final_;
    for (int s = 0; s < S; ++s) {
        const double *row = A + (size_t)s * W + 1;
        int any = 0;
        for (int j = 0; j < S; ++j) any |= (row[j] != 0.0);
        for (int d = 0; d < dim; ++d) {
            /* strided reduce over the selected columns: strictly left to right */
            double acc = 0.0; int first = 1;
            for (int j = 0; j < S; ++j) {
                if (any && row[j] == 0.0) continue;   /* mask = all-ones when no coefficient is set */
                double term = w->K[(size_t)j * dim + d] * row[j];
                acc = first ? term : acc + term;
                first = 0;
            }
            w->tmp[d] = h * acc;                       /* intermediate_dstate */
            w->ystage[d] = y[d] + w->tmp[d];
        }
        orc_rhs(prob, prm, t + h * A[(size_t)s * W], w->ystage, w->K + (size_t)s * dim);
    }
    if (fsal) {
        /* integrator_types.py:303-305: dState = last stage's increment, final_rhs = last stage */
        for (int d = 0; d < dim; ++d) { w->dS[d] = w->tmp[d]; w->f_final[d] = w->K[(size_t)(S - 1) * dim + d]; }
    } else {
        for (int d = 0; d < dim; ++d) {
            for (int j = 0; j < S; ++j) w->terms[j] = w->K[(size_t)j * dim + d] * B[1 + j];
            w->dS[d] = h * np_sum_contig(w->terms, S);
        }
        for (int d = 0; d < dim; ++d) w->ystage[d] = y[d] + w->dS[d];
        orc_rhs(prob, prm, t + h, w->ystage, w->f_final);
    }
    if (tab->final_rows == 2) {
        for (int d = 0; d < dim; ++d) {
            for (int j = 0; j < S; ++j) w->terms[j] = (B[1 + j] - B[W + 1 + j]) * w->K[(size_t)j * dim + d];
            w->diff[d] = h * np_sum_contig(w->terms, S);
        }
    } else {
        for (int d = 0; d < dim; ++d) w->diff[d] = 0.0;
    }
}

static int g_factor_ulps = 0;
void orc_set_factor_perturbation(int ulps) { g_factor_ulps = ulps; }

/* update_timestep (integrator_template.py:46-93).  attempt_no = 0 for the first attempt of a
 * step.  eps_last/eps_last_last carry the retry-chain history.  Returns the factor. */
static double erk_controller(erk_work *w, const double *y, double h, double rtol, double atol,
                             double order, int attempt_no, double *eps_last, double *eps_last_last)
{
    const int dim = w->dim;
    for (int d = 0; d < dim; ++d) {
        double sc = np_maximum(fabs(y[d]), fabs(w->dS[d] / h));
        w->scale[d] = attempt_no == 0 ? sc : 0.8 * w->scale[d] + 0.2 * sc;
        w->tmp[d] = w->diff[d] / (atol + rtol * w->scale[d]);
    }
    double eps_cur = 1.0 / l2_norm(w->tmp, dim);
    double corr;
    if (attempt_no == 0) {
        corr = eps_cur > 0.0 ? pow(eps_cur, 1.0 / order) : 1.0;
        *eps_last = eps_cur;
    } else if (attempt_no == 1) {
        corr = eps_cur > 0.0 ? pow(eps_cur, 1.0 / order) : 1.0;
        corr = corr * (*eps_last > 0.0 ? pow(*eps_last, 1.0 / order) : 1.0);
        *eps_last_last = *eps_last; *eps_last = eps_cur;
    } else {
        double k1 = pow(eps_cur, 0.55 / order);
        double k2 = pow(*eps_last, -0.27 / order);
        double k3 = pow(*eps_last_last, 0.05 / order);
        corr = k1 > 0.0 ? k1 : 1.0;
        corr = corr * (k2 > 0.0 ? k2 : 1.0);
        corr = corr * (k3 > 0.0 ? k3 : 1.0);
        *eps_last_last = *eps_last; *eps_last = eps_cur;
    }
    double factor = 1.0 + atan(0.8 * corr - 1.0);
    /* test knob (orc_set_factor_perturbation): nudge the step-size factor by a few ulps -- the size of the
     * difference between two correctly-behaving libm's -- to measure how far chaos carries it */
    for (int u = 0; u < (g_factor_ulps < 0 ? -g_factor_ulps : g_factor_ulps); ++u)
        factor = nextafter(factor, g_factor_ulps > 0 ? INFINITY : -INFINITY);
    return factor;
}

static int tableau_is_fsal(const orc_rk_tableau *tab)
{
    const int S = tab->stages, W = S + 1;
    for (int j = 1; j < W; ++j)
        if (tab->inter[(size_t)(S - 1) * W + j] != tab->final_[j]) return 0;
    return 1;
}

static void dense_record(orc_dense *dn, int dim, double t0, double t1, const double *y0,
                         const double *y1, const double *f0, const double *f1)
{
    if (dn->count < dn->capacity) {
        size_t k = (size_t)dn->count;
        dn->t0[k] = t0; dn->t1[k] = t1;
        memcpy(dn->y0 + k * dim, y0, sizeof(double) * dim);
        memcpy(dn->y1 + k * dim, y1, sizeof(double) * dim);
        memcpy(dn->f0 + k * dim, f0, sizeof(double) * dim);
        memcpy(dn->f1 + k * dim, f1, sizeof(double) * dim);
    }
    dn->count++;
}

int orc_erk_trajectory(const orc_problem *prob, const orc_rk_tableau *tab, const double *prm,
                       double *y, double *t_io, double *dt_io, double tf, double rtol, double atol,
                       int64_t max_steps, int64_t *accepted, int64_t *attempts, orc_dense *dense)
{
    const int dim = prob->dim;
    const int adaptive = tab->final_rows == 2;
    const int fsal = tableau_is_fsal(tab);
    const double reject_below = 0.9 * 0.9;          /* 0.9**2, integrator_template.py:93 */
    double t = *t_io, dt = *dt_io;
    int64_t acc = 0, att = 0;
    int status = ORC_STATUS_OK;
    erk_work w;
    if (!erk_work_alloc(&w, dim, tab->stages)) { erk_work_free(&w); return -1; }
    double *ynew = malloc(sizeof(double) * dim);

    /* differential_system.py:956-957,971-974 */
    if (fabs(tf - t) < ORC_EPS4) goto done;
    if ((dt > 0) != (tf - t > 0) && dt != 0.0) dt = -dt;
    if (fabs(dt) > fabs(tf - t)) { dt = fabs(tf - t) * 0.5; if (tf - t < 0) dt = -dt; }

    while (dt != 0.0 && fabs(tf - t) >= ORC_EPS32) {                        /* :996 */
        if (max_steps >= 0 && acc >= max_steps) { status = ORC_STATUS_RUNNING; break; }
        int final_step = fabs(dt + t) > fabs(tf);                             /* :997 */
        double h = final_step ? (tf - t) : dt, h0 = h, new_dt = h;
        double eps_last = 0.0, eps_last_last = 0.0;
        int redo = 0;
        /* integrator_types.py:168-173: initial_rhs for dense output */
        if (w.have_final) memcpy(w.f_init, w.f_final, sizeof(double) * dim);
        else orc_rhs(prob, prm, t, y, w.f_init);
        for (int trial = 0; trial <= ORC_RETRIES; ++trial) {
            att++;
            erk_attempt(prob, tab, prm, &w, t, y, h, fsal);
            w.have_final = 1;
            if (!adaptive) { new_dt = h; redo = 0; break; }
            double f = erk_controller(&w, y, h, rtol, atol, tab->order, trial, &eps_last, &eps_last_last);
            new_dt = f * h;
            redo = f < reject_below;
            if (!redo) break;
            /* integrator_types.py:206: min(timestep, current_timestep); sign-aware for tf<t0 */
            h = h0 >= 0 ? fmin(new_dt, h0) : -fmin(fabs(new_dt), fabs(h0));
        }
        if (redo) { status = ORC_STATUS_TOL_FAIL; break; }                  /* :220-224 */
        for (int d = 0; d < dim; ++d) ynew[d] = y[d] + w.dS[d];              /* :1015 */
        if (dense) dense_record(dense, dim, t, t + h, y, ynew, w.f_init, w.f_final);
        memcpy(y, ynew, sizeof(double) * dim);
        t = t + h;                                                            /* :1016 */
        acc++;
        if (!final_step) dt = new_dt;                                         /* :1070-1071 */
    }
done:
    *t_io = t; *dt_io = dt;
    if (accepted) *accepted = acc;
    if (attempts) *attempts = att;
    free(ynew);
    erk_work_free(&w);
    return status;
}

/* ----------------------------------------------------------- symplectic ---- */

int orc_symplectic_trajectory(const orc_problem *prob, int S, const double *T, const double *prm,
                              double *y, double *t_io, double *dt_io, double tf,
                              int64_t max_steps, int64_t *steps_out)
{
    const int dim = prob->dim, half = prob->drift_count >= 0 ? prob->drift_count : dim / 2;
    double t = *t_io, dt = *dt_io;
    int64_t steps = 0;
    int status = ORC_STATUS_OK;
    double *dS = calloc(dim, sizeof(double)), *ys = malloc(sizeof(double) * dim), *f = malloc(sizeof(double) * dim);

    if (fabs(tf - t) < ORC_EPS4) goto done;
    if ((dt > 0) != (tf - t > 0) && dt != 0.0) dt = -dt;
    if (fabs(dt) > fabs(tf - t)) { dt = fabs(tf - t) * 0.5; if (tf - t < 0) dt = -dt; }

    while (dt != 0.0 && fabs(tf - t) >= ORC_EPS32) {
        if (max_steps >= 0 && steps >= max_steps) { status = ORC_STATUS_RUNNING; break; }
        int final_step = fabs(dt + t) > fabs(tf);
        double h = final_step ? (tf - t) : dt;
        double tau = t;
        for (int d = 0; d < dim; ++d) dS[d] = dS[d] * 0.0;                  /* integrator_types.py:404 */
        for (int s = 0; s < S; ++s) {
            for (int d = 0; d < dim; ++d) ys[d] = y[d] + dS[d];
            orc_rhs(prob, prm, tau, ys, f);
            tau = tau + h * T[3 * s + 1];                                    /* :412 (column 1) */
            /* kick mask = second half of axis 0 (:359-362); masks are exact 0/1 */
            double cd = T[3 * s + 1] * 1.0 + T[3 * s + 2] * 0.0;
            double ck = T[3 * s + 1] * 0.0 + T[3 * s + 2] * 1.0;
            for (int d = 0; d < dim; ++d) {
                double aux = h * f[d];
                dS[d] = dS[d] + aux * (d < half ? cd : ck);                   /* :413 */
            }
        }
        for (int d = 0; d < dim; ++d) y[d] = y[d] + dS[d];
        t = t + h;
        steps++;
    }
done:
    *t_io = t; *dt_io = dt;
    if (steps_out) *steps_out = steps;
    free(dS); free(ys); free(f);
    return status;
}

/* ------------------------------------------------------ ensemble drivers ---- */

void orc_erk_ensemble(const orc_problem *prob, const orc_rk_tableau *tab, const double *params,
                      double *y, double *t, double *dt, int64_t n, int64_t first, int64_t count,
                      double tf, double rtol, double atol,
                      int32_t *accepted, int32_t *rejected, int32_t *status)
{
    const int dim = prob->dim, np_ = prob->n_params;
    double *yl = malloc(sizeof(double) * dim), *pl = malloc(sizeof(double) * (np_ > 0 ? np_ : 1));
    for (int64_t i = first; i < first + count && i < n; ++i) {
        for (int d = 0; d < dim; ++d) yl[d] = y[(size_t)d * n + i];
        for (int p = 0; p < np_; ++p) pl[p] = params[(size_t)p * n + i];
        int64_t acc = 0, att = 0;
        int st = orc_erk_trajectory(prob, tab, pl, yl, &t[i], &dt[i], tf, rtol, atol, -1, &acc, &att, NULL);
        for (int d = 0; d < dim; ++d) y[(size_t)d * n + i] = yl[d];
        if (accepted) accepted[i] = (int32_t)acc;
        if (rejected) rejected[i] = (int32_t)(att - acc);
        if (status) status[i] = st;
    }
    free(yl); free(pl);
}

void orc_symplectic_ensemble(const orc_problem *prob, int S, const double *T, const double *params,
                             double *y, double *t, double *dt, int64_t n, int64_t first, int64_t count,
                             double tf, int32_t *steps, int32_t *status)
{
    const int dim = prob->dim, np_ = prob->n_params;
    double *yl = malloc(sizeof(double) * dim), *pl = malloc(sizeof(double) * (np_ > 0 ? np_ : 1));
    for (int64_t i = first; i < first + count && i < n; ++i) {
        for (int d = 0; d < dim; ++d) yl[d] = y[(size_t)d * n + i];
        for (int p = 0; p < np_; ++p) pl[p] = params[(size_t)p * n + i];
        int64_t st_n = 0;
        int st = orc_symplectic_trajectory(prob, S, T, pl, yl, &t[i], &dt[i], tf, -1, &st_n);
        for (int d = 0; d < dim; ++d) y[(size_t)d * n + i] = yl[d];
        if (steps) steps[i] = (int32_t)st_n;
        if (status) status[i] = st;
    }
    free(yl); free(pl);
}

/* --------------------------------------------------------- dense output ---- */

void orc_hermite_eval(int dim, double t0, double t1, const double *p0, const double *p1,
                      const double *m0, const double *m1, double tq, double *out)
{
    double trange = t1 - t0;
    double x = (tq - t0) / trange;
    if (x == 0.0) { memcpy(out, p0, sizeof(double) * dim); return; }
    if (x == 1.0) { memcpy(out, p1, sizeof(double) * dim); return; }
    double x2 = x * x, x3 = x2 * x;
    double h00 = 2 * x3 - 3 * x2 + 1;
    double h10 = x3 - 2 * x2 + x;
    double h01 = -2 * x3 + 3 * x2;
    double h11 = x3 - x2;
    for (int d = 0; d < dim; ++d)
        out[d] = ((h00 * p0[d] + h10 * trange * m0[d]) + h01 * p1[d]) + h11 * trange * m1[d];
}

void orc_dense_eval(const orc_dense *dn, int dim, double tq, double *out)
{
    /* first recorded step whose END time is >= tq, clamped (differential_system.py:205-215) */
    int64_t n = dn->count < dn->capacity ? dn->count : dn->capacity;
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (dn->t1[mid] >= tq) hi = mid; else lo = mid + 1;
    }
    size_t k = (size_t)lo;
    orc_hermite_eval(dim, dn->t0[k], dn->t1[k], dn->y0 + k * dim, dn->y1 + k * dim,
                     dn->f0 + k * dim, dn->f1 + k * dim, tq, out);
}
