// erk_stage.cu -- K2: stage-level explicit Runge-Kutta kernels for right-hand sides that are NOT fused
// (an arbitrary torch callable, or the dense linear RHS evaluated by a GEMM).  Per attempt the host side
// issues, on one stream and without synchronising:
//     dsb_erk_stage_begin                      step start: final-step clamp, h, h0 (differential_system.py:997-1002)
//     for s in stages: dsb_erk_stage_state(s)  y_stage = y + h * sum_j a_sj K_j, t_stage = t + c_s h
//                      K[s] = rhs(t_stage, y_stage)            <- the user's code, between our kernels
//     dsb_erk_stage_finish                     dY, error estimate, per-trajectory L2 norm, controller,
//                                              masked commit, counters, active count
// Any explicit tableau with up to DSB_MAX_STAGES stages (RK14(12) has 35).  Layout: [dim][n], trajectory index
// fastest, every array; K[s] are separate [dim][n] tensors (the tensors the user's RHS returned).
#include <math_constants.h>

#include "common.cuh"
#include "fastmath.cuh"
#include "plan.h"

namespace dsb {

constexpr int FLAG_ACTIVE = 1, FLAG_FINAL = 2, FLAG_ACCEPTED = 4;
constexpr int FLAG_FINAL_LEG = 8;      // re-integration to the root of a terminal event in progress (events are off)
constexpr int EV_NE = DSB_MAX_STAGE_EVENTS;  // ev_state rows: 0 t_prev, 1 dt_prev, 2.. last occurrence per event, 2+NE.. g per event

__device__ __forceinline__ double tf_of(const dsb_erk_stage_args &a, long long i) { return a.tf_per ? a.tf_per[i] : a.tf; }

// g_e(t, y + dy) of a hyperplane event for trajectory i (dy = nullptr: g_e(t, y))
__device__ __forceinline__ double stage_ev_g(const dsb_erk_stage_args &a, int e, long long i, double t, const double *dy)
{
    const double *w = a.ev_coef + (long long)e * (a.dim + 2);
    double g = w[a.dim] * t - w[a.dim + 1];
    for (int d = 0; d < a.dim; ++d) {
        if (w[d] != 0.0) {
            const long long q = (long long)d * a.n + i;
            g = fma(w[d], dy ? a.y[q] + dy[q] : a.y[q], g);
        }
    }
    return g;
}

struct StageRow {
    int nterms;
    int idx[DSB_MAX_STAGES];
    double coef[DSB_MAX_STAGES];
    double c;
};

struct FinalRows {
    int stages, adaptive, fsal, order_p;
    double order;
    double b[DSB_MAX_STAGES], db[DSB_MAX_STAGES];
};

struct KPtrs {
    const double *k[DSB_MAX_STAGES];
};

// ---- begin: integrate()-entry logic on the first call, step-start logic on every attempt -------------------
__global__ void stage_begin_kernel(dsb_erk_stage_args a, int first_call)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int fl = a.flags[i];
    double t = a.t[i], dt = a.dt[i];
    if (first_call) {
        int st0 = a.status ? a.status[i] : (int)DSB_TRAJ_DONE;
        bool go = st0 != (int)DSB_TRAJ_TOL_FAIL && st0 != (int)DSB_TRAJ_EVENT;
        double span = tf_of(a, i) - t;
        if (go) {
            if (fabs(span) < DSB_EPS4) go = false;                                  // differential_system.py:956
            else {
                if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;            // :971
                if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);   // :973-974
            }
        }
        go = go && (dt != 0.0) && (fabs(tf_of(a, i) - t) >= DSB_EPS32);             // :996
        fl = go ? FLAG_ACTIVE : 0;
        a.trial[i] = 0;
        a.dt[i] = dt;
        const bool frozen = st0 == (int)DSB_TRAJ_TOL_FAIL || st0 == (int)DSB_TRAJ_EVENT;
        if (a.status) a.status[i] = go ? (int)DSB_TRAJ_RUNNING : (frozen ? st0 : (int)DSB_TRAJ_DONE);
        if (a.n_events > 0) {
            for (int e = 0; e < a.n_events; ++e) {
                a.ev_state[(long long)(2 + e) * a.n + i] = CUDART_NAN;
                a.ev_state[(long long)(2 + EV_NE + e) * a.n + i] = stage_ev_g(a, e, i, t, nullptr);
            }
        }
    }
    if (i == 0) a.counters[1] = 0;                 // active count, re-accumulated by the controller kernel
    fl &= ~FLAG_ACCEPTED;
    if ((fl & FLAG_ACTIVE) && a.trial[i] == 0) {
        const double tfi = tf_of(a, i);
        bool final_step = fabs(dt + t) > fabs(tfi);                                 // :997
        double h = final_step ? (tfi - t) : dt;
        a.h[i] = h;
        a.h0[i] = h;
        fl = final_step ? (fl | FLAG_FINAL) : (fl & ~FLAG_FINAL);
    }
    a.flags[i] = fl;
}

// ---- stage state: non-zero coefficients only, left to right (runge_kutta_methods.py:45-48) -----------------
template <bool STRICT>
__global__ void stage_state_kernel(dsb_erk_stage_args a, KPtrs K, StageRow row, int keep_increment)
{
    using A = Ar<STRICT>;
    const long long total = (long long)a.dim * a.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % a.n;
        const double h = a.h[i];
        double sum = 0.0;
        for (int q = 0; q < row.nterms; ++q) sum = A::mad(K.k[row.idx[q]][e], row.coef[q], sum);
        const double inc = A::mul(h, sum);
        a.y_stage[e] = STRICT ? A::add(a.y[e], inc) : fma(h, sum, a.y[e]);
        if (keep_increment) a.dY[e] = inc;       // FSAL: dState = last stage's increment (integrator_types.py:303-305)
        if (e < a.n) a.t_stage[i] = A::add(a.t[i], A::mul(h, row.c));
    }
}

// Ensemble-shaped variant (n >= 64): thread = VEC adjacent trajectories (128-bit accesses for VEC = 2), grid.y
// walks the state rows; no integer division per element, h is loaded once per thread.
template <int VEC> struct Pack;
template <> struct Pack<1> { double v[1]; };
template <> struct __align__(16) Pack<2> { double v[2]; };

template <bool STRICT, int VEC>
__global__ void stage_state_rows_kernel(dsb_erk_stage_args a, KPtrs K, StageRow row, int keep_increment)
{
    using A = Ar<STRICT>;
    using P = Pack<VEC>;
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= a.n) return;
    const P h = *reinterpret_cast<const P *>(a.h + col);
    if (blockIdx.y == 0) {
        const P t = *reinterpret_cast<const P *>(a.t + col);
        P ts;
#pragma unroll
        for (int v = 0; v < VEC; ++v) ts.v[v] = A::add(t.v[v], A::mul(h.v[v], row.c));
        *reinterpret_cast<P *>(a.t_stage + col) = ts;
    }
    for (int d = blockIdx.y; d < a.dim; d += gridDim.y) {
        const long long e = (long long)d * a.n + col;
        P sum;
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum.v[v] = 0.0;
        for (int q = 0; q < row.nterms; ++q) {
            const P kv = *reinterpret_cast<const P *>(K.k[row.idx[q]] + e);
#pragma unroll
            for (int v = 0; v < VEC; ++v) sum.v[v] = A::mad(kv.v[v], row.coef[q], sum.v[v]);
        }
        const P y = *reinterpret_cast<const P *>(a.y + e);
        P ys, inc;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            inc.v[v] = A::mul(h.v[v], sum.v[v]);
            ys.v[v] = STRICT ? A::add(y.v[v], inc.v[v]) : fma(h.v[v], sum.v[v], y.v[v]);
        }
        *reinterpret_cast<P *>(a.y_stage + e) = ys;
        if (keep_increment) *reinterpret_cast<P *>(a.dY + e) = inc;
    }
}

template <bool STRICT, int VEC>
__global__ void stage_commit_rows_kernel(dsb_erk_stage_args a)
{
    using A = Ar<STRICT>;
    using P = Pack<VEC>;
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (col >= a.n) return;
    bool acc[VEC], any = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) { acc[v] = a.flags[col + v] & FLAG_ACCEPTED; any |= acc[v]; }
    if (!any) return;
    for (int d = blockIdx.y; d < a.dim; d += gridDim.y) {
        const long long e = (long long)d * a.n + col;
        P y = *reinterpret_cast<const P *>(a.y + e);
        const P dy = *reinterpret_cast<const P *>(a.dY + e);
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (acc[v]) y.v[v] = A::add(y.v[v], dy.v[v]);
        *reinterpret_cast<P *>(a.y + e) = y;
    }
}

// ---- finish 1/3: increment, error estimate, scale, partial squared norms -----------------------------------
// grid.x tiles trajectories, grid.y tiles the state dimension into `chunks`; thread (i, chunk) walks its d-range.
template <bool STRICT>
__global__ void stage_increment_kernel(dsb_erk_stage_args a, KPtrs K, FinalRows fr, int d_per_chunk)
{
    using A = Ar<STRICT>;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int chunk = blockIdx.y;
    const int d0 = chunk * d_per_chunk, d1 = min(a.dim, d0 + d_per_chunk);
    const double h = a.h[i];
    const int trial = a.trial[i];
    const int S = fr.stages;
    double x = 0.0;
    bool first = true;
    for (int d = d0; d < d1; ++d) {
        const long long e = (long long)d * a.n + i;
        double sb, se;
        if (STRICT) {
            // numpy contiguous reduce over ALL stages: < 8 left to right, else 8 accumulators + tail
            double r[8], rb = 0.0, re = 0.0;
            double q[8];
            if (S < 8) {
                for (int j = 0; j < S; ++j) {
                    double kv = K.k[j][e];
                    double tb = __dmul_rn(kv, fr.b[j]), te = __dmul_rn(fr.db[j], kv);
                    rb = j == 0 ? tb : __dadd_rn(rb, tb);
                    re = j == 0 ? te : __dadd_rn(re, te);
                }
            } else {
                const int full = S - (S % 8);
                for (int j = 0; j < 8; ++j) { double kv = K.k[j][e]; r[j] = __dmul_rn(kv, fr.b[j]); q[j] = __dmul_rn(fr.db[j], kv); }
                for (int j0 = 8; j0 < full; j0 += 8)
                    for (int j = 0; j < 8; ++j) {
                        double kv = K.k[j0 + j][e];
                        r[j] = __dadd_rn(r[j], __dmul_rn(kv, fr.b[j0 + j]));
                        q[j] = __dadd_rn(q[j], __dmul_rn(fr.db[j0 + j], kv));
                    }
                rb = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                re = __dadd_rn(__dadd_rn(__dadd_rn(q[0], q[1]), __dadd_rn(q[2], q[3])), __dadd_rn(__dadd_rn(q[4], q[5]), __dadd_rn(q[6], q[7])));
                for (int j = full; j < S; ++j) {
                    double kv = K.k[j][e];
                    rb = __dadd_rn(rb, __dmul_rn(kv, fr.b[j]));
                    re = __dadd_rn(re, __dmul_rn(fr.db[j], kv));
                }
            }
            sb = rb; se = re;
        } else {
            sb = 0.0; se = 0.0;
            for (int j = 0; j < S; ++j) {
                double kv = K.k[j][e];
                sb = fma(kv, fr.b[j], sb);
                se = fma(kv, fr.db[j], se);
            }
        }
        // FSAL: dState is the last stage's increment, kept by the state kernel (integrator_types.py:303-305)
        double dY = fr.fsal ? a.dY[e] : A::mul(h, sb);
        double err = A::mul(h, se);
        a.dY[e] = dY;
        if (fr.adaptive) {
            double slope = STRICT ? dY / h : (fr.fsal ? fm::fast_div(dY, h) : sb);
            double sc = np_maximum(fabs(a.y[e]), fabs(slope));
            if (trial > 0) sc = A::add(A::mul(0.8, a.scale[e]), A::mul(0.2, sc));
            a.scale[e] = sc;
            // array-valued tolerances (reference integrator_types.py:37-38 accepts them): one (rtol, atol) per component
            double tol = A::mad(a.rtol_vec ? a.rtol_vec[d] : a.rtol, sc, a.atol_vec ? a.atol_vec[d] : a.atol);
            double qv = STRICT ? err / tol : fm::fast_div(err, tol);
            x = first ? __dmul_rn(qv, qv) : fma(qv, qv, x);
            first = false;
        }
    }
    a.partial[(long long)chunk * a.n + i] = x;
}

// Same as stage_increment_kernel for FEW trajectories with a LARGE state (the reference's plain use: one big
// system, one dt): a CTA owns one (chunk, trajectory) pair, its threads stride over the chunk's state elements and
// the partial squared norm is reduced in a fixed order (warp shuffles, then warp 0 over the warp sums), so the
// result is deterministic.  The summation order differs from the FMA chain of the narrow kernel at rounding level.
template <bool STRICT>
__global__ void stage_increment_wide_kernel(dsb_erk_stage_args a, KPtrs K, FinalRows fr, int d_per_chunk)
{
    using A = Ar<STRICT>;
    const long long i = blockIdx.y;
    const int chunk = blockIdx.x;
    const int d0 = chunk * d_per_chunk, d1 = min(a.dim, d0 + d_per_chunk);
    const double h = a.h[i];
    const int trial = a.trial[i];
    const int S = fr.stages;
    double x = 0.0;
    for (int d = d0 + threadIdx.x; d < d1; d += blockDim.x) {
        const long long e = (long long)d * a.n + i;
        double sb = 0.0, se = 0.0;
        for (int j = 0; j < S; ++j) {
            const double kv = K.k[j][e];
            sb = A::mad(kv, fr.b[j], sb);
            se = A::mad(kv, fr.db[j], se);
        }
        const double dY = fr.fsal ? a.dY[e] : A::mul(h, sb);
        const double err = A::mul(h, se);
        a.dY[e] = dY;
        if (fr.adaptive) {
            double sc = np_maximum(fabs(a.y[e]), fabs(dY / h));
            if (trial > 0) sc = A::add(A::mul(0.8, a.scale[e]), A::mul(0.2, sc));
            a.scale[e] = sc;
            const double qv = err / A::mad(a.rtol_vec ? a.rtol_vec[d] : a.rtol, sc, a.atol_vec ? a.atol_vec[d] : a.atol);
            x = fma(qv, qv, x);
        }
    }
    __shared__ double warp_sum[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) a.partial[(long long)chunk * a.n + i] = v;
    }
}

__device__ __forceinline__ double guarded_pow2(double base, double expo)
{
    double k = pow(base, expo);
    return k > 0.0 ? k : 1.0;
}

// ---- finish 2/3: per-trajectory controller + bookkeeping (integrator_template.py:46-93, types :188-224) ----
template <bool STRICT>
__global__ void stage_controller_kernel(dsb_erk_stage_args a, FinalRows fr, int chunks)
{
    using A = Ar<STRICT>;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int still_active = 0;
    if (i < a.n) {
        int fl = a.flags[i];
        if (fl & FLAG_ACTIVE) {
            const double h = a.h[i], h0 = a.h0[i];
            int trial = a.trial[i];
            bool redo = false;
            double new_dt = h;
            if (fr.adaptive) {
                double x = a.partial[i];
                for (int c = 1; c < chunks; ++c) x = __dadd_rn(x, a.partial[(long long)c * a.n + i]);
                double *hist = a.hist;                                  // [0] x_last [1] x_last_last [2] corr_first
                const double x_last = hist[i], x_ll = hist[a.n + i], corr_first = hist[2 * a.n + i];
                double corr;
                if (trial < 2) {
                    double c_cur;
                    if (STRICT) { double eps = 1.0 / sqrt(x); c_cur = eps > 0.0 ? pow(eps, 1.0 / fr.order) : 1.0; }
                    else if (x > 1e-280 && x < 1e280 && fr.order_p > 0) c_cur = fm::inv_root_rt(x, fr.order_p);
                    else if (x == 0.0) c_cur = CUDART_INF;
                    else if (!(x < CUDART_INF)) c_cur = 1.0;
                    else c_cur = pow(x, -0.5 / fr.order);
                    corr = trial == 0 ? c_cur : A::mul(c_cur, corr_first);
                    if (trial == 0) hist[2 * a.n + i] = c_cur;
                } else {
                    double k1 = guarded_pow2(1.0 / sqrt(x), 0.55 / fr.order);
                    double k2 = guarded_pow2(1.0 / sqrt(x_last), -0.27 / fr.order);
                    double k3 = guarded_pow2(1.0 / sqrt(x_ll), 0.05 / fr.order);
                    corr = A::mul(A::mul(k1, k2), k3);
                }
                hist[a.n + i] = x_last;
                hist[i] = x;
                double f = A::add(1.0, atan(A::sub(A::mul(0.8, corr), 1.0)));
                new_dt = A::mul(f, h);
                redo = f < DSB_REJECT_BELOW;
            }
            if (!redo) {
                double dt = a.dt[i];
                if (a.n_events > 0) {                                   // the roll-back of a terminal event needs these
                    a.ev_state[i] = a.t[i];
                    a.ev_state[a.n + i] = dt;
                }
                double t = A::add(a.t[i], h);
                a.t[i] = t;
                if (!(fl & FLAG_FINAL)) { dt = new_dt; a.dt[i] = dt; }
                a.trial[i] = 0;
                if (a.accepted) a.accepted[i] += 1;
                fl |= FLAG_ACCEPTED;
                bool more = (dt != 0.0) && (fabs(tf_of(a, i) - t) >= DSB_EPS32);
                if (!more) {
                    fl &= ~FLAG_ACTIVE;
                    if (a.status) a.status[i] = (fl & FLAG_FINAL_LEG) ? (int)DSB_TRAJ_EVENT : (int)DSB_TRAJ_DONE;
                }
                else if (a.one_step) { fl &= ~FLAG_ACTIVE; if (a.status) a.status[i] = (int)DSB_TRAJ_RUNNING; }
            } else {
                if (a.rejected) a.rejected[i] += 1;
                trial += 1;
                if (trial > DSB_MAX_RETRIES) {
                    fl &= ~FLAG_ACTIVE;
                    a.trial[i] = 0;
                    if (a.status) a.status[i] = (int)DSB_TRAJ_TOL_FAIL;
                } else {
                    a.trial[i] = trial;
                    a.h[i] = h0 >= 0.0 ? fmin(new_dt, h0) : -fmin(fabs(new_dt), fabs(h0));
                }
            }
            a.flags[i] = fl;
            still_active = (fl & FLAG_ACTIVE) ? 1 : 0;
        }
    }
    // block-reduced count of trajectories that still need attempts -> counters[1]
    int cnt = __syncthreads_count(still_active);
    if (threadIdx.x == 0 && cnt) atomicAdd((unsigned long long *)&a.counters[1], (unsigned long long)cnt);
}

// ---- finish 3/3: masked commit y += dY where the attempt was accepted (differential_system.py:1015) --------
template <bool STRICT>
__global__ void stage_commit_kernel(dsb_erk_stage_args a)
{
    using A = Ar<STRICT>;
    const long long total = (long long)a.dim * a.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % a.n;
        if (a.flags[i] & FLAG_ACCEPTED) a.y[e] = A::add(a.y[e], a.dY[e]);
    }
}

// ---- events, phase 0 tail: y_stage = y + dY, t_stage = t (already advanced where the attempt was accepted) ----
__global__ void stage_y1_kernel(dsb_erk_stage_args a)
{
    const long long total = (long long)a.dim * a.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        a.y_stage[e] = a.y[e] + a.dY[e];
        if (e < a.n) a.t_stage[e] = a.t[e];
    }
}

// ---- events, phase 1: locate / record / terminal roll-back, one thread per trajectory.  The same algorithm as the
// event block of the fused kernel (erk_fused.cuh): differential_system.py:29-57, 60-167, 1020-1066. -------------------
// ext_roots != nullptr: the roots were located by the caller (dsb_event_search_*, any event callable): [n_events][n]
// event times, NaN = no event of that kind in this step; the polynomial solve is skipped.
__global__ void stage_events_kernel(dsb_erk_stage_args a, const double *__restrict__ k0, const double *__restrict__ f1,
                                    const double *__restrict__ ext_roots)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int fl = a.flags[i];
    if (!(fl & FLAG_ACCEPTED) || (fl & FLAG_FINAL_LEG)) return;
    const long long n = a.n;
    const int D = a.dim, W = a.dim + 2, ne = a.n_events;
    const double t1 = a.t[i], t0 = a.ev_state[i], dt_prev = a.ev_state[n + i], trange = t1 - t0;
    // Along the step's interpolant an affine event g(t, y, dy) = wy.y + wdy.dy + wt t - c is the polynomial
    //   G(tau) = [Hermite cubic of (gy0, gy1, a, b)](tau) + wdy . grad(tau),
    // with gy = wy.y + wt t - c at the ends, a = trange (wy.m0 + wt), b = trange (wy.m1 + wt), and
    //   wdy . grad(tau) = M0 + tau (-6 (P - Q)/trange - 4 M0 - 2 M1) + tau^2 (6 (P - Q)/trange + 3 M0 + 3 M1)
    // (utilities/interpolation.py:63-78; P, Q = wdy.p0, wdy.p1; M0, M1 = wdy.m0, wdy.m1).  g at the step start is
    // formed from the step's own data (p0, m0 = K[0]) -- what the reference's root finder evaluates at t_prev.
    double pc[EV_NE][4];                      // G(tau) = pc0 + pc1 tau + pc2 tau^2 + pc3 tau^3
    unsigned mask = 0;
    if (ext_roots)
        for (int e = 0; e < ne; ++e) {
            const double r = ext_roots[(long long)e * n + i];
            if (r == r) mask |= 1u << e;
        }
    for (int e = 0; e < ne && !ext_roots; ++e) {
        const double *w = a.ev_coef + (long long)e * W;
        const double *wd = a.ev_coef_dy ? a.ev_coef_dy + (long long)e * D : nullptr;
        double gy0 = w[D] * t0 - w[D + 1], gy1 = w[D] * t1 - w[D + 1], dg0 = w[D], dg1 = w[D];
        double P = 0.0, Q = 0.0, M0 = 0.0, M1 = 0.0;
        bool has_dy = false;
        for (int d = 0; d < D; ++d) {
            const long long q = (long long)d * n + i;
            const bool uy = w[d] != 0.0, ud = wd && wd[d] != 0.0;
            if (!uy && !ud) continue;
            const double p0 = a.y[q], p1 = a.y[q] + a.dY[q], m0 = k0[q], m1 = f1[q];
            if (uy) { gy0 = fma(w[d], p0, gy0); gy1 = fma(w[d], p1, gy1); dg0 = fma(w[d], m0, dg0); dg1 = fma(w[d], m1, dg1); }
            if (ud) { P = fma(wd[d], p0, P); Q = fma(wd[d], p1, Q); M0 = fma(wd[d], m0, M0); M1 = fma(wd[d], m1, M1); has_dy = true; }
        }
        const double ca = trange * dg0, cb = trange * dg1;
        pc[e][0] = gy0;
        pc[e][1] = ca;
        pc[e][2] = 3.0 * (gy1 - gy0) - 2.0 * ca - cb;
        pc[e][3] = 2.0 * (gy0 - gy1) + ca + cb;
        double g0 = gy0, g1 = gy1;
        if (has_dy) {
            const double s6 = 6.0 * (P - Q) / trange;
            pc[e][0] += M0;
            pc[e][1] += -s6 - 4.0 * M0 - 2.0 * M1;
            pc[e][2] += s6 + 3.0 * M0 + 3.0 * M1;
            g0 += M0;
            g1 += M1;
        }
        a.ev_state[(long long)(2 + EV_NE + e) * n + i] = g1;
        const int dir = a.ev_flags[e] & 3;
        const bool rising = g0 < 0.0;
        if (g0 * g1 < 0.0 && (dir == 0 || (dir == 1) == rising)) mask |= 1u << e;
    }
    if (!mask) return;
    double root[EV_NE];
    int which[EV_NE], cnt = 0;
    for (int e = 0; e < ne; ++e) {
        if (!(mask >> e & 1u)) continue;
        if (ext_roots) {
            root[cnt] = ext_roots[(long long)e * n + i];
            which[cnt] = e;
            cnt++;
            continue;
        }
        const double c0 = pc[e][0], c1 = pc[e][1], c2 = pc[e][2], c3 = pc[e][3];
        const double g0 = c0, g1 = ((c3 + c2) + c1) + c0;
        double lo = 0.0, hi = 1.0, tau = g0 / (g0 - g1);
        for (int it = 0; it < 64; ++it) {
            const double G = fma(fma(fma(c3, tau, c2), tau, c1), tau, c0);
            if (G == 0.0) break;
            if ((G < 0.0) == (g0 < 0.0)) lo = tau; else hi = tau;
            const double dG = fma(fma(3.0 * c3, tau, 2.0 * c2), tau, c1);
            const double step = G / dG;
            if (fabs(step) <= 2.3e-16) break;
            double nxt = tau - step;
            if (!(nxt > lo && nxt < hi)) nxt = 0.5 * (lo + hi);
            tau = nxt;
            if (hi - lo <= 4.0e-16) break;
        }
        root[cnt] = fma(tau, trange, t0);
        which[cnt] = e;
        cnt++;
    }
    for (int p = 1; p < cnt; ++p)
        for (int q = p; q > 0 && (trange >= 0.0 ? root[q] < root[q - 1] : root[q] > root[q - 1]); --q) {
            double tr = root[q]; root[q] = root[q - 1]; root[q - 1] = tr;
            int tw = which[q]; which[q] = which[q - 1]; which[q - 1] = tw;
        }
    bool term = false;
    double term_root = 0.0;
    for (int p = 0; p < cnt && !term; ++p) {
        const int e = which[p];
        const double rt = root[p], last = a.ev_state[(long long)(2 + e) * n + i];
        if (!(last == last) || fabs(rt - last) > 2.9e-11) {
            a.ev_state[(long long)(2 + e) * n + i] = rt;
            const int c = a.ev_count[i];
            if (c < a.ev_capacity) {
                double *rec = a.ev_records + ((long long)i * a.ev_capacity + c) * W;
                const double ta = (rt - t0) / trange;
                rec[0] = rt;
                rec[1 + D] = (double)e;
                const double t2 = __dmul_rn(ta, ta), t3 = __dmul_rn(t2, ta);
                const double h00 = __dadd_rn(__dsub_rn(__dmul_rn(2.0, t3), __dmul_rn(3.0, t2)), 1.0);
                const double h10 = __dadd_rn(__dsub_rn(t3, __dmul_rn(2.0, t2)), ta);
                const double h01 = __dadd_rn(__dmul_rn(-2.0, t3), __dmul_rn(3.0, t2));
                const double h11 = __dsub_rn(t3, t2);
                for (int d = 0; d < D; ++d) {
                    const long long q = (long long)d * n + i;
                    const double p0 = a.y[q], p1 = a.y[q] + a.dY[q];
                    rec[1 + d] = ta == 0.0 ? p0 : (ta == 1.0 ? p1 :
                        __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h00, p0), __dmul_rn(__dmul_rn(h10, trange), k0[q])),
                                            __dmul_rn(h01, p1)), __dmul_rn(__dmul_rn(h11, trange), f1[q])));
                }
            }
            a.ev_count[i] = c + 1;
        }
        if (a.ev_flags[e] & 4) { term = true; term_root = rt; }
    }
    if (term) {
        // drop the step and integrate to the root (differential_system.py:1051-1055, entry clamp :973-974)
        const bool was_active = fl & FLAG_ACTIVE;
        a.t[i] = t0;
        if (a.accepted) a.accepted[i] -= 1;
        a.tf_per[i] = term_root;
        fl &= ~(FLAG_ACCEPTED | FLAG_FINAL);
        const double span = term_root - t0;
        double dt = dt_prev;
        bool active;
        if (fabs(span) < DSB_EPS4) {
            active = false;
            if (a.status) a.status[i] = (int)DSB_TRAJ_EVENT;
        } else {
            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
            active = (dt != 0.0) && (fabs(span) >= DSB_EPS32);
            if (a.status) a.status[i] = active ? (int)DSB_TRAJ_RUNNING : (int)DSB_TRAJ_EVENT;
        }
        a.dt[i] = dt;
        a.trial[i] = 0;
        fl = active ? (fl | FLAG_ACTIVE | FLAG_FINAL_LEG) : ((fl & ~FLAG_ACTIVE) | FLAG_FINAL_LEG);
        a.flags[i] = fl;
        if (active && !was_active) atomicAdd((unsigned long long *)&a.counters[1], 1ull);
        if (!active && was_active) atomicAdd((unsigned long long *)&a.counters[1], ~0ull);      // -1
    }
}

static void build_final(const dsb_erk_plan *p, FinalRows &fr);

// the event kernel over any buffers that follow the dsb_erk_stage_args conventions (used by symplectic.cu as well)
int launch_stage_events(const dsb_erk_stage_args &view, const double *k0, const double *f1, const double *ext_roots,
                        cudaStream_t stream)
{
    stage_events_kernel<<<(unsigned)((view.n + 127) / 128), 128, 0, stream>>>(view, k0, f1, ext_roots);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

static void build_rows(const dsb_erk_plan *p, int s, StageRow &row)
{
    const int S = p->stages, W = S + 1;
    row.nterms = 0;
    row.c = p->inter[s * W];
    for (int j = 0; j < S; ++j) {
        double v = p->inter[s * W + 1 + j];
        if (v != 0.0) { row.idx[row.nterms] = j; row.coef[row.nterms] = v; row.nterms++; }
    }
}

static void build_final(const dsb_erk_plan *p, FinalRows &fr)
{
    const int S = p->stages, W = S + 1;
    fr.stages = S;
    fr.adaptive = p->final_rows == 2;
    fr.order = p->order;
    double p2 = 2.0 * p->order;
    fr.order_p = (p2 == (double)(int)p2 && p2 >= 1.0 && p2 <= 64.0) ? (int)p2 : 0;
    fr.fsal = 1;
    for (int j = 0; j < S; ++j) {
        fr.b[j] = p->fin[1 + j];
        fr.db[j] = fr.adaptive ? p->fin[1 + j] - p->fin[W + 1 + j] : 0.0;
        if (p->inter[(S - 1) * W + 1 + j] != fr.b[j]) fr.fsal = 0;
    }
}

static int grid_for(long long total, int threads, int sms)
{
    long long blocks = (total + threads - 1) / threads;
    long long cap = (long long)sms * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// rows of the state per grid.y slot: enough CTAs to fill the machine (~16 per SM), each walking many rows so that
// the per-CTA setup (parameter structs, h) is amortised
static unsigned rows_grid_y(unsigned gx, int dim, int sms)
{
    unsigned want = ((unsigned)sms * 16u + gx - 1) / gx;
    if (want < 1) want = 1;
    return want < (unsigned)dim ? want : (unsigned)dim;
}

static int check_args(const dsb_erk_plan *h, const dsb_erk_stage_args *a, const char *who)
{
    if (!h || !a) { set_error("%s: null handle or args", who); return DSB_ERR_BAD_ARGUMENT; }
    if (h->stages > DSB_MAX_STAGES) { set_error("%s: %d stages exceed DSB_MAX_STAGES", who, h->stages); return DSB_ERR_UNSUPPORTED; }
    if (a->n < 0 || a->dim < 1 || !a->y || !a->t || !a->dt || !a->h || !a->h0 || !a->y_stage || !a->t_stage || !a->dY ||
        !a->scale || !a->hist || !a->partial || !a->trial || !a->flags || !a->counters || a->chunks < 1) {
        set_error("%s: missing buffer", who);
        return DSB_ERR_BAD_ARGUMENT;
    }
    return DSB_OK;
}

int stage_begin(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int first_call, cudaStream_t stream)
{
    int rc = check_args(h, a, "dsb_erk_stage_begin");
    if (rc != DSB_OK || a->n == 0) return rc;
    stage_begin_kernel<<<(unsigned)((a->n + 127) / 128), 128, 0, stream>>>(*a, first_call);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

// 128-bit accesses (VEC = 2) need every array 16-byte aligned: torch allocations are, but a right-hand side may return
// a view at an odd 8-byte offset.
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static int pick_vec(const dsb_erk_plan *h, const dsb_erk_stage_args *a)
{
    if (a->n % 2 != 0) return 1;
    bool ok = aligned16(a->y) && aligned16(a->t) && aligned16(a->h) && aligned16(a->dY) && aligned16(a->y_stage) &&
              aligned16(a->t_stage) && aligned16(a->flags);
    for (int j = 0; j < h->stages && ok; ++j) ok = !a->K[j] || aligned16(a->K[j]);
    return ok ? 2 : 1;
}

int stage_state(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int stage, cudaStream_t stream)
{
    int rc = check_args(h, a, "dsb_erk_stage_state");
    if (rc != DSB_OK || a->n == 0) return rc;
    if (stage < 0 || stage >= h->stages) { set_error("dsb_erk_stage_state: stage %d out of range", stage); return DSB_ERR_BAD_ARGUMENT; }
    StageRow row;
    build_rows(h, stage, row);
    KPtrs K;
    for (int j = 0; j < DSB_MAX_STAGES; ++j) K.k[j] = j < h->stages ? a->K[j] : nullptr;
    for (int q = 0; q < row.nterms; ++q)
        if (!K.k[row.idx[q]]) { set_error("dsb_erk_stage_state: K[%d] is null", row.idx[q]); return DSB_ERR_BAD_ARGUMENT; }
    const long long total = (long long)a->dim * a->n;
    const int grid = grid_for(total, 256, h->sm_count);
    FinalRows fr;
    build_final(h, fr);
    const int keep = (fr.fsal && stage == h->stages - 1) ? 1 : 0;
    const bool strict = h->flags & DSB_FLAG_STRICT;
    if (a->n >= 64) {
        const int vec = pick_vec(h, a);
        const unsigned gx = (unsigned)((a->n / vec + 127) / 128);
        dim3 g(gx, rows_grid_y(gx, a->dim, h->sm_count));
        if (vec == 2) {
            if (strict) stage_state_rows_kernel<true, 2><<<g, 128, 0, stream>>>(*a, K, row, keep);
            else stage_state_rows_kernel<false, 2><<<g, 128, 0, stream>>>(*a, K, row, keep);
        } else {
            if (strict) stage_state_rows_kernel<true, 1><<<g, 128, 0, stream>>>(*a, K, row, keep);
            else stage_state_rows_kernel<false, 1><<<g, 128, 0, stream>>>(*a, K, row, keep);
        }
    } else if (strict) stage_state_kernel<true><<<grid, 256, 0, stream>>>(*a, K, row, keep);
    else stage_state_kernel<false><<<grid, 256, 0, stream>>>(*a, K, row, keep);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

// parts: 1 = increment + error norm + controller, 2 = masked commit, 3 = both
static int stage_finish_parts(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int parts, cudaStream_t stream)
{
    int rc = check_args(h, a, "dsb_erk_stage_finish");
    if (rc != DSB_OK || a->n == 0) return rc;
    FinalRows fr;
    build_final(h, fr);
    KPtrs K;
    for (int j = 0; j < DSB_MAX_STAGES; ++j) {
        K.k[j] = j < h->stages ? a->K[j] : nullptr;
        if (j < h->stages && !K.k[j]) { set_error("dsb_erk_stage_finish: K[%d] is null", j); return DSB_ERR_BAD_ARGUMENT; }
    }
    const bool strict = h->flags & DSB_FLAG_STRICT;
    const int chunks = a->chunks;
    const int d_per_chunk = (a->dim + chunks - 1) / chunks;
    dim3 g1((unsigned)((a->n + 127) / 128), (unsigned)chunks);
    const unsigned g2 = (unsigned)((a->n + 127) / 128);
    const int g3 = grid_for((long long)a->dim * a->n, 256, h->sm_count);
    const bool rows = a->n >= 64;
    const int vec = pick_vec(h, a);
    const unsigned grx = (unsigned)((a->n / vec + 127) / 128);
    dim3 gr(grx, rows_grid_y(grx, a->dim, h->sm_count));
    // few trajectories, large state (e.g. ensemble=False with a big system): parallelise over the state instead
    const bool wide = a->n < 32 && a->dim >= 8192;
    if (wide && (parts & 1)) {
        dim3 gw((unsigned)chunks, (unsigned)a->n);
        if (strict) stage_increment_wide_kernel<true><<<gw, 256, 0, stream>>>(*a, K, fr, d_per_chunk);
        else stage_increment_wide_kernel<false><<<gw, 256, 0, stream>>>(*a, K, fr, d_per_chunk);
    }
    if (strict) {
        if (parts & 1) {
            if (!wide) stage_increment_kernel<true><<<g1, 128, 0, stream>>>(*a, K, fr, d_per_chunk);
            stage_controller_kernel<true><<<g2, 128, 0, stream>>>(*a, fr, chunks);
        }
        if (parts & 2) {
            if (rows && vec == 2) stage_commit_rows_kernel<true, 2><<<gr, 128, 0, stream>>>(*a);
            else if (rows) stage_commit_rows_kernel<true, 1><<<gr, 128, 0, stream>>>(*a);
            else stage_commit_kernel<true><<<g3, 256, 0, stream>>>(*a);
        }
    } else {
        if (parts & 1) {
            if (!wide) stage_increment_kernel<false><<<g1, 128, 0, stream>>>(*a, K, fr, d_per_chunk);
            stage_controller_kernel<false><<<g2, 128, 0, stream>>>(*a, fr, chunks);
        }
        if (parts & 2) {
            if (rows && vec == 2) stage_commit_rows_kernel<false, 2><<<gr, 128, 0, stream>>>(*a);
            else if (rows) stage_commit_rows_kernel<false, 1><<<gr, 128, 0, stream>>>(*a);
            else stage_commit_kernel<false><<<g3, 256, 0, stream>>>(*a);
        }
    }
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int stage_finish(const dsb_erk_plan *h, const dsb_erk_stage_args *a, cudaStream_t stream)
{
    if (a && a->n_events > 0) {
        set_error("dsb_erk_stage_finish: with events the attempt ends with dsb_erk_stage_finish_events (phases 0 and 1)");
        return DSB_ERR_BAD_ARGUMENT;
    }
    return stage_finish_parts(h, a, 3, stream);
}

int stage_finish_events(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int phase, const double *f1, const double *roots,
                        cudaStream_t stream)
{
    int rc = check_args(h, a, "dsb_erk_stage_finish_events");
    if (rc != DSB_OK) return rc;
    if (a->n_events < 1 || a->n_events > DSB_MAX_STAGE_EVENTS || !a->ev_coef || !a->ev_flags || !a->ev_records || !a->ev_count ||
        !a->ev_state || !a->tf_per || a->ev_capacity < 1 || (phase != 0 && phase != 1) || (phase == 1 && !f1) || !a->K[0]) {
        set_error("dsb_erk_stage_finish_events: needs 1..%d events with ev_coef, ev_flags, ev_records, ev_count, ev_state, "
                  "tf_per, ev_capacity >= 1, K[0], phase 0 or 1, and f1 in phase 1", DSB_MAX_STAGE_EVENTS);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (a->n == 0) return DSB_OK;
    if (phase == 0) {
        rc = stage_finish_parts(h, a, 1, stream);
        if (rc != DSB_OK) return rc;
        stage_y1_kernel<<<grid_for((long long)a->dim * a->n, 256, h->sm_count), 256, 0, stream>>>(*a);
    } else {
        stage_events_kernel<<<(unsigned)((a->n + 127) / 128), 128, 0, stream>>>(*a, a->K[0], f1, roots);
        DSB_CUDA_CHECK(cudaGetLastError());
        return stage_finish_parts(h, a, 2, stream);
    }
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // namespace dsb
