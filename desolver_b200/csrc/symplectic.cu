// symplectic.cu -- K4: explicit symplectic splitting integrators (BABs9o7H, ABAs5o6H, symplectic Euler).
//
//  (A) nbody_symplectic_kernel: fused, one launch per integrate(): the whole fixed-step loop of
//      OdeSystem.integrate (differential_system.py:996-1002,1015-1018: last step truncated to land on tf)
//      around ExplicitSymplecticIntegrator.step (integrators/integrator_types.py:402-417) with the softened
//      gravitational N-body right-hand side (desolver_b200/rhs_plugins.py:nbody) evaluated in shared memory.
//      One thread per body, `SPB` systems per CTA; all-pairs forces, j summed left to right like the
//      reference's sum over axis -2.  The reference evaluates the full RHS at all S stages and multiplies the
//      unused half by a 0 mask; here a force evaluation happens only at stages whose kick
//      coefficient is non-zero (10 of 19 for BABs9o7H) -- identical results for finite states.
//  (B) stage-level kernels for any other right-hand side (user torch callable between them).
// Kick (momentum) variables = second half of axis 0 of the state, drift (position) = first half
// (integrator_types.py:359-362).  Time advances by tableau column 1 (:412).
#include <math_constants.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace dsb {

struct SympTableau {
    int stages;
    double drift[DSB_MAX_STAGES];   // column 1
    double kick[DSB_MAX_STAGES];    // column 2
};

}  // namespace dsb

struct dsb_symp_plan {
    int device = 0, rhs_id = 0, dim = 0, drift_count = 0, sm_count = 0;
    unsigned flags = 0;
    dsb::SympTableau tab;
};

namespace dsb {

// x^(-3/2) for normal positive x in 6 FP64 instructions: with the MUFU.RSQ64H seed s ~ x^(-1/2) and e = 1 - x s^2
// (|e| < 2^-21), x^(-3/2) = s^3 (1 - e)^(-3/2) = s^3 (1 + 3/2 e + 15/8 e^2 + O(e^3)); the dropped term is below 1e-19.
// (CUDA's rsqrt() carries a denormal / inf side branch worth five issue slots per pair interaction, and cubing a corrected
// 1/sqrt costs one instruction more: 30.5 -> 28.5 ms for config 3.)
__device__ __forceinline__ double inv_r3_pos(double x)
{
    double s;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x));
    const double s2 = s * s;
    const double e = fma(-x, s2, 1.0);
    return (s2 * s) * fma(fma(e, 1.875, 1.5), e, 1.0);
}

// ------------------------------------------------------------------------------------------ (A) fused N-body
template <bool STRICT>
__global__ void nbody_symplectic_kernel(dsb_symp_run_args a, SympTableau T)
{
    using A = Ar<STRICT>;
    extern __shared__ double4 smem4[];
    const int nb = a.bodies;
    const int spb = blockDim.y;
    // stage positions and masses of this CTA's systems as (x, y, z, m): two LDS.128 per partner body
    double4 *sq = smem4 + (size_t)threadIdx.y * nb;
    const long long sys = (long long)blockIdx.x * spb + threadIdx.y;
    const int i = threadIdx.x;
    const bool valid = sys < a.n;
    const long long n = a.n;
    const long long s_ = valid ? sys : 0;
    const double my_mass = a.masses[i];

    double q0[3], v0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q0[k] = a.y[(long long)(i * 3 + k) * n + s_];
        v0[k] = a.y[(long long)(nb * 3 + i * 3 + k) * n + s_];
    }
    double t = a.t[s_], dt = a.dt[s_];
    const double tf = a.tf;
    int steps = 0;
    bool go = valid;
    if (a.first_call) {                                                     // differential_system.py:956-957,971-974
        double span = tf - t;
        if (fabs(span) < DSB_EPS4) go = false;
        else {
            if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;
            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
        }
    }

    // the loop condition is uniform within a system (same t, dt in all its threads); systems of one CTA may
    // stop at different steps, so every thread keeps hitting the barriers until the whole CTA is done
    for (;;) {
        bool more = go && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32) && (a.max_steps < 0 || steps < a.max_steps);
        if (!__syncthreads_or(more)) break;
        const bool final_step = fabs(dt + t) > fabs(tf);                    // :997
        const double h = final_step ? (tf - t) : dt;
        double dq[3] = {0.0, 0.0, 0.0}, dv[3] = {0.0, 0.0, 0.0};
        for (int s = 0; s < T.stages; ++s) {
            const double c1 = T.drift[s], c2 = T.kick[s];
            double vcur[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) vcur[k] = A::add(v0[k], dv[k]);
            if (c2 != 0.0) {                                                // kick stage: forces at q + dS_q
                const double qx = A::add(q0[0], dq[0]), qy = A::add(q0[1], dq[1]), qz = A::add(q0[2], dq[2]);
                sq[i] = make_double4(qx, qy, qz, my_mass);
                __syncthreads();
                double ax = 0.0, ay = 0.0, az = 0.0;
                // The partner loop includes j == i: there d = 0 and the term is 0 * (m r2^-1.5) = +0 for any
                // softening eps2 > 0 -- exactly what the reference's w * (1 - eye) yields (for eps2 = 0 both give NaN).
#pragma unroll 4
                for (int j = 0; j < nb; ++j) {
                    const double4 pj = sq[j];
                    const double dx = A::sub(pj.x, qx), dy = A::sub(pj.y, qy), dz = A::sub(pj.z, qz);
                    double w;
                    if (STRICT) {
                        double r2 = A::add(A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz)), a.eps2);
                        w = A::mul(pj.w, pow(r2, -1.5));
                        if (j == i) w = A::mul(w, 0.0);
                    } else {
                        double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, a.eps2)));
                        w = pj.w * inv_r3_pos(r2);
                    }
                    ax = A::mad(dx, w, ax);
                    ay = A::mad(dy, w, ay);
                    az = A::mad(dz, w, az);
                }
                __syncthreads();
                dv[0] = A::add(dv[0], A::mul(A::mul(h, A::mul(a.G, ax)), c2));     // dS += (h*f) * coef  (:409-413)
                dv[1] = A::add(dv[1], A::mul(A::mul(h, A::mul(a.G, ay)), c2));
                dv[2] = A::add(dv[2], A::mul(A::mul(h, A::mul(a.G, az)), c2));
            }
            if (c1 != 0.0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) dq[k] = A::add(dq[k], A::mul(A::mul(h, vcur[k]), c1));
            }
        }
        if (more) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { q0[k] = A::add(q0[k], dq[k]); v0[k] = A::add(v0[k], dv[k]); }
            t = A::add(t, h);
            steps++;
        }
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            a.y[(long long)(i * 3 + k) * n + sys] = q0[k];
            a.y[(long long)(nb * 3 + i * 3 + k) * n + sys] = v0[k];
        }
        if (i == 0) {
            a.t[sys] = t;
            a.dt[sys] = dt;
            if (a.steps) a.steps[sys] += steps;
            bool more = go && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32);
            if (a.status) a.status[sys] = more ? (int)DSB_TRAJ_RUNNING : (int)DSB_TRAJ_DONE;
        }
    }
}

#ifndef DSB_NBODY_X2_MINB
#define DSB_NBODY_X2_MINB 3
#endif
// ------------------------------------------------------------------------------------------ (A2) two bodies per thread
// The same loop with bodies i and i + nb/2 in ONE thread (nb even): every partner (x, y, z, m) read from shared memory
// serves two interactions, which halves the LDS.128 / address / loop instructions per pair (7 -> ~4 non-FP64 issue slots
// next to the 16 FP64 ones).  Each body's arithmetic -- operations and the left-to-right order over j -- is that of the
// kernel above, so the results are identical bit for bit.
template <bool STRICT>
__global__ void __launch_bounds__(128, DSB_NBODY_X2_MINB) nbody_symplectic_kernel_x2(dsb_symp_run_args a, SympTableau T)
{
    using A = Ar<STRICT>;
    extern __shared__ double4 smem4[];
    const int nb = a.bodies, half = nb >> 1;
    const int spb = blockDim.y;
    double4 *sq = smem4 + (size_t)threadIdx.y * nb;
    const long long sys = (long long)blockIdx.x * spb + threadIdx.y;
    const int i0 = threadIdx.x, i1 = threadIdx.x + half;
    const bool valid = sys < a.n;
    const long long n = a.n;
    const long long s_ = valid ? sys : 0;
    const double m0 = a.masses[i0], m1 = a.masses[i1];

    double q0[2][3], v0[2][3];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int i = b ? i1 : i0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            q0[b][k] = a.y[(long long)(i * 3 + k) * n + s_];
            v0[b][k] = a.y[(long long)(nb * 3 + i * 3 + k) * n + s_];
        }
    }
    double t = a.t[s_], dt = a.dt[s_];
    const double tf = a.tf;
    int steps = 0;
    bool go = valid;
    if (a.first_call) {
        double span = tf - t;
        if (fabs(span) < DSB_EPS4) go = false;
        else {
            if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;
            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
        }
    }
    for (;;) {
        bool more = go && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32) && (a.max_steps < 0 || steps < a.max_steps);
        if (!__syncthreads_or(more)) break;
        const bool final_step = fabs(dt + t) > fabs(tf);
        const double h = final_step ? (tf - t) : dt;
        double dq[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, dv[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        for (int s = 0; s < T.stages; ++s) {
            const double c1 = T.drift[s], c2 = T.kick[s];
            double vcur[2][3];
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 3; ++k) vcur[b][k] = A::add(v0[b][k], dv[b][k]);
            if (c2 != 0.0) {
                const double ax0 = A::add(q0[0][0], dq[0][0]), ay0 = A::add(q0[0][1], dq[0][1]), az0 = A::add(q0[0][2], dq[0][2]);
                const double ax1 = A::add(q0[1][0], dq[1][0]), ay1 = A::add(q0[1][1], dq[1][1]), az1 = A::add(q0[1][2], dq[1][2]);
                sq[i0] = make_double4(ax0, ay0, az0, m0);
                sq[i1] = make_double4(ax1, ay1, az1, m1);
                __syncthreads();
                double fx0 = 0.0, fy0 = 0.0, fz0 = 0.0, fx1 = 0.0, fy1 = 0.0, fz1 = 0.0;
#pragma unroll 4
                for (int j = 0; j < nb; ++j) {
                    const double4 pj = sq[j];
                    const double dx0 = A::sub(pj.x, ax0), dy0 = A::sub(pj.y, ay0), dz0 = A::sub(pj.z, az0);
                    const double dx1 = A::sub(pj.x, ax1), dy1 = A::sub(pj.y, ay1), dz1 = A::sub(pj.z, az1);
                    double w0, w1;
                    if (STRICT) {
                        double r0 = A::add(A::add(A::add(A::mul(dx0, dx0), A::mul(dy0, dy0)), A::mul(dz0, dz0)), a.eps2);
                        double r1 = A::add(A::add(A::add(A::mul(dx1, dx1), A::mul(dy1, dy1)), A::mul(dz1, dz1)), a.eps2);
                        w0 = A::mul(pj.w, pow(r0, -1.5));
                        w1 = A::mul(pj.w, pow(r1, -1.5));
                        if (j == i0) w0 = A::mul(w0, 0.0);
                        if (j == i1) w1 = A::mul(w1, 0.0);
                    } else {
                        double r0 = fma(dz0, dz0, fma(dy0, dy0, fma(dx0, dx0, a.eps2)));
                        double r1 = fma(dz1, dz1, fma(dy1, dy1, fma(dx1, dx1, a.eps2)));
                        w0 = pj.w * inv_r3_pos(r0);
                        w1 = pj.w * inv_r3_pos(r1);
                    }
                    fx0 = A::mad(dx0, w0, fx0); fy0 = A::mad(dy0, w0, fy0); fz0 = A::mad(dz0, w0, fz0);
                    fx1 = A::mad(dx1, w1, fx1); fy1 = A::mad(dy1, w1, fy1); fz1 = A::mad(dz1, w1, fz1);
                }
                __syncthreads();
                dv[0][0] = A::add(dv[0][0], A::mul(A::mul(h, A::mul(a.G, fx0)), c2));
                dv[0][1] = A::add(dv[0][1], A::mul(A::mul(h, A::mul(a.G, fy0)), c2));
                dv[0][2] = A::add(dv[0][2], A::mul(A::mul(h, A::mul(a.G, fz0)), c2));
                dv[1][0] = A::add(dv[1][0], A::mul(A::mul(h, A::mul(a.G, fx1)), c2));
                dv[1][1] = A::add(dv[1][1], A::mul(A::mul(h, A::mul(a.G, fy1)), c2));
                dv[1][2] = A::add(dv[1][2], A::mul(A::mul(h, A::mul(a.G, fz1)), c2));
            }
            if (c1 != 0.0) {
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int k = 0; k < 3; ++k) dq[b][k] = A::add(dq[b][k], A::mul(A::mul(h, vcur[b][k]), c1));
            }
        }
        if (more) {
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int k = 0; k < 3; ++k) { q0[b][k] = A::add(q0[b][k], dq[b][k]); v0[b][k] = A::add(v0[b][k], dv[b][k]); }
            t = A::add(t, h);
            steps++;
        }
    }
    if (valid) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = b ? i1 : i0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a.y[(long long)(i * 3 + k) * n + sys] = q0[b][k];
                a.y[(long long)(nb * 3 + i * 3 + k) * n + sys] = v0[b][k];
            }
        }
        if (i0 == 0) {
            a.t[sys] = t;
            a.dt[sys] = dt;
            if (a.steps) a.steps[sys] += steps;
            bool more = go && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32);
            if (a.status) a.status[sys] = more ? (int)DSB_TRAJ_RUNNING : (int)DSB_TRAJ_DONE;
        }
    }
}

// ------------------------------------------------------------------------------------- (B) stage-level kernels
constexpr int SFLAG_ACTIVE = 1;
constexpr int SFLAG_STEPPED = 4;       // took a step in the commit just done (mask of dsb_dense_append); cleared by begin
constexpr int SFLAG_FINAL_LEG = 8;     // re-integration to the root of a terminal event (same bit as in erk_stage.cu)

int launch_stage_events(const dsb_erk_stage_args &view, const double *k0, const double *f1, const double *ext_roots,
                        cudaStream_t stream);

__device__ __forceinline__ double symp_tf(const dsb_symp_stage_args &a, long long i) { return a.tf_per ? a.tf_per[i] : a.tf; }

__global__ void symp_begin_kernel(dsb_symp_stage_args a, int first_call)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double t = a.t[i], dt = a.dt[i];
    int fl = a.flags[i];
    const double tfi = symp_tf(a, i);
    if (first_call) {
        bool go = true;
        double span = tfi - t;
        if (a.n_events > 0) {
            fl &= ~SFLAG_FINAL_LEG;
            for (int e = 0; e < a.n_events; ++e) {
                const double *w = a.ev_coef + (long long)e * (a.dim + 2);
                double g = w[a.dim] * t - w[a.dim + 1];
                for (int d = 0; d < a.dim; ++d)
                    if (w[d] != 0.0) g = fma(w[d], a.y[(long long)d * a.n + i], g);
                a.ev_state[(long long)(2 + e) * a.n + i] = CUDART_NAN;
                a.ev_state[(long long)(2 + DSB_MAX_STAGE_EVENTS + e) * a.n + i] = g;
            }
        }
        if (fabs(span) < DSB_EPS4) go = false;
        else {
            if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;
            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
        }
        a.dt[i] = dt;
        fl = (fl & SFLAG_FINAL_LEG) | (go ? SFLAG_ACTIVE : 0);
    }
    bool more = (fl & SFLAG_ACTIVE) && (dt != 0.0) && (fabs(tfi - t) >= DSB_EPS32);
    fl = (fl & SFLAG_FINAL_LEG) | (more ? SFLAG_ACTIVE : 0);
    a.flags[i] = fl;
    if (a.status) a.status[i] = more ? (int)DSB_TRAJ_RUNNING : ((fl & SFLAG_FINAL_LEG) ? (int)DSB_TRAJ_EVENT : (int)DSB_TRAJ_DONE);
    bool final_step = fabs(dt + t) > fabs(tfi);
    a.h[i] = more ? (final_step ? (tfi - t) : dt) : 0.0;
    a.t_stage[i] = t;
    if (i == 0) a.counters[1] = 0;
}

// dS = 0 and y_stage = y at the start of a step
__global__ void symp_reset_kernel(dsb_symp_stage_args a)
{
    const long long total = (long long)a.dim * a.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        a.dS[e] = 0.0;
        a.y_stage[e] = a.y[e];
    }
}

// dS += (h * f) * (drift or kick coefficient); y_stage = y + dS; t_stage += h * drift   (:407-413)
template <bool STRICT>
__global__ void symp_accumulate_kernel(dsb_symp_stage_args a, const double *f, double c_drift, double c_kick, int drift_count)
{
    using A = Ar<STRICT>;
    const long long total = (long long)a.dim * a.n;
    const long long half = (long long)drift_count * a.n;       // leading part of axis 0 = drift variables
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % a.n;
        const double h = a.h[i];
        const double coef = e < half ? c_drift : c_kick;
        double ds = a.dS[e];
        if (coef != 0.0) ds = A::add(ds, A::mul(A::mul(h, f[e]), coef));
        a.dS[e] = ds;
        a.y_stage[e] = A::add(a.y[e], ds);
        if (e < a.n) a.t_stage[i] = A::add(a.t_stage[i], A::mul(h, c_drift));
    }
}

template <bool STRICT, int MASK = SFLAG_ACTIVE>
__global__ void symp_commit_kernel(dsb_symp_stage_args a)
{
    using A = Ar<STRICT>;
    const long long total = (long long)a.dim * a.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % a.n;
        if (a.flags[i] & MASK) a.y[e] = A::add(a.y[e], a.dS[e]);
    }
}

__global__ void symp_advance_kernel(dsb_symp_stage_args a)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int still = 0;
    if (i < a.n && (a.flags[i] & SFLAG_ACTIVE)) {
        const int leg = a.flags[i] & SFLAG_FINAL_LEG;
        double dt = a.dt[i];
        if (a.n_events > 0) {                                        // the roll-back of a terminal event needs these
            a.ev_state[i] = a.t[i];
            a.ev_state[a.n + i] = dt;
        }
        double t = __dadd_rn(a.t[i], a.h[i]);
        a.t[i] = t;
        if (a.steps) a.steps[i] += 1;
        still = (dt != 0.0) && (fabs(symp_tf(a, i) - t) >= DSB_EPS32);
        a.flags[i] = (still ? SFLAG_ACTIVE : 0) | SFLAG_STEPPED | leg;
        if (!still && a.status) a.status[i] = leg ? (int)DSB_TRAJ_EVENT : (int)DSB_TRAJ_DONE;
    }
    int cnt = __syncthreads_count(still);
    if (threadIdx.x == 0 && cnt) atomicAdd((unsigned long long *)&a.counters[1], (unsigned long long)cnt);
}

static int grid_for(long long total, int threads, int sms)
{
    long long blocks = (total + threads - 1) / threads;
    long long cap = (long long)sms * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace dsb

using namespace dsb;

extern "C" {

int dsb_symp_create(dsb_symp_handle *out, int device, int rhs_id, int dim, int drift_count, int stages,
                    const double *tableau, unsigned flags)
{
    if (!out || !tableau || stages < 1 || stages > DSB_MAX_STAGES || dim < 2 || drift_count < 0 || drift_count > dim) {
        set_error("dsb_symp_create: bad argument (stages=%d dim=%d drift_count=%d; stages <= %d)", stages, dim, drift_count,
                  DSB_MAX_STAGES);
        return DSB_ERR_BAD_ARGUMENT;
    }
    DeviceGuard guard(device);                 // the caller's current device is restored on return
    int sm_count = 0;
    DSB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    dsb_symp_plan *p = new dsb_symp_plan;
    p->device = device; p->rhs_id = rhs_id; p->dim = dim; p->drift_count = drift_count; p->flags = flags;
    p->tab.stages = stages;
    for (int s = 0; s < stages; ++s) { p->tab.drift[s] = tableau[3 * s + 1]; p->tab.kick[s] = tableau[3 * s + 2]; }
    p->sm_count = sm_count;
    *out = p;
    return DSB_OK;
}

int dsb_symp_destroy(dsb_symp_handle h) { delete h; return DSB_OK; }

int dsb_symp_run(dsb_symp_handle h, const dsb_symp_run_args *a, void *stream)
{
    if (!h || !a || !a->y || !a->t || !a->dt) { set_error("dsb_symp_run: null argument"); return DSB_ERR_BAD_ARGUMENT; }
    if (h->rhs_id != DSB_RHS_NBODY) { set_error("dsb_symp_run: fused symplectic kernel exists for the N-body right-hand side only"); return DSB_ERR_UNSUPPORTED; }
    if (a->bodies < 1 || a->bodies > 1024 || h->dim != 6 * a->bodies || h->drift_count != 3 * a->bodies || !a->masses) {
        set_error("dsb_symp_run: state dim %d does not match 2*bodies*3 with bodies=%d", h->dim, a->bodies);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (a->n == 0) return DSB_OK;
    DeviceGuard guard(h->device);
    // two bodies per thread where the body count allows it (DSB_NBODY_X2=0 keeps one body per thread)
    static const bool x2_on = [] { const char *e = getenv("DSB_NBODY_X2"); return !e || atoi(e) != 0; }();
    if (x2_on && a->bodies % 2 == 0 && a->bodies >= 8 && a->bodies <= 256) {
        const int tpb = a->bodies / 2;
        int spb2 = 128 / tpb;
        if (spb2 < 1) spb2 = 1;
        dim3 block2((unsigned)tpb, (unsigned)spb2);
        unsigned grid2 = (unsigned)((a->n + spb2 - 1) / spb2);
        size_t smem2 = sizeof(double4) * (size_t)spb2 * a->bodies;
        if (h->flags & DSB_FLAG_STRICT) nbody_symplectic_kernel_x2<true><<<grid2, block2, smem2, (cudaStream_t)stream>>>(*a, h->tab);
        else nbody_symplectic_kernel_x2<false><<<grid2, block2, smem2, (cudaStream_t)stream>>>(*a, h->tab);
        DSB_CUDA_CHECK(cudaGetLastError());
        return DSB_OK;
    }
    int spb = 128 / a->bodies;
    if (spb < 1) spb = 1;
    dim3 block((unsigned)a->bodies, (unsigned)spb);
    unsigned grid = (unsigned)((a->n + spb - 1) / spb);
    size_t smem = sizeof(double4) * (size_t)spb * a->bodies;
    if (h->flags & DSB_FLAG_STRICT) nbody_symplectic_kernel<true><<<grid, block, smem, (cudaStream_t)stream>>>(*a, h->tab);
    else nbody_symplectic_kernel<false><<<grid, block, smem, (cudaStream_t)stream>>>(*a, h->tab);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

static int symp_check(dsb_symp_handle h, const dsb_symp_stage_args *a, const char *who)
{
    if (!h || !a || !a->y || !a->t || !a->dt || !a->h || !a->dS || !a->y_stage || !a->t_stage || !a->flags || !a->counters ||
        a->dim != h->dim || a->n < 0) {
        set_error("%s: bad argument", who);
        return DSB_ERR_BAD_ARGUMENT;
    }
    return DSB_OK;
}

int dsb_symp_stage_begin(dsb_symp_handle h, const dsb_symp_stage_args *a, int first_call, void *stream)
{
    int rc = symp_check(h, a, "dsb_symp_stage_begin");
    if (rc != DSB_OK || a->n == 0) return rc;
    DeviceGuard guard(h->device);
    symp_begin_kernel<<<(unsigned)((a->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*a, first_call);
    symp_reset_kernel<<<grid_for((long long)a->dim * a->n, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(*a);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_symp_stage_accumulate(dsb_symp_handle h, const dsb_symp_stage_args *a, int stage, const double *f, void *stream)
{
    int rc = symp_check(h, a, "dsb_symp_stage_accumulate");
    if (rc != DSB_OK || a->n == 0) return rc;
    DeviceGuard guard(h->device);
    if (stage < 0 || stage >= h->tab.stages || !f) { set_error("dsb_symp_stage_accumulate: bad stage or null f"); return DSB_ERR_BAD_ARGUMENT; }
    const int grid = grid_for((long long)a->dim * a->n, 256, h->sm_count);
    if (h->flags & DSB_FLAG_STRICT)
        symp_accumulate_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, f, h->tab.drift[stage], h->tab.kick[stage], h->drift_count);
    else
        symp_accumulate_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, f, h->tab.drift[stage], h->tab.kick[stage], h->drift_count);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_symp_stage_commit(dsb_symp_handle h, const dsb_symp_stage_args *a, void *stream)
{
    int rc = symp_check(h, a, "dsb_symp_stage_commit");
    if (rc != DSB_OK || a->n == 0) return rc;
    DeviceGuard guard(h->device);
    const int grid = grid_for((long long)a->dim * a->n, 256, h->sm_count);
    if (h->flags & DSB_FLAG_STRICT) symp_commit_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    else symp_commit_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    symp_advance_kernel<<<(unsigned)((a->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*a);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

static int symp_commit_events(dsb_symp_handle h, const dsb_symp_stage_args *a, int phase, const double *f0, const double *f1,
                              const double *roots, void *stream);

int dsb_symp_stage_commit_events(dsb_symp_handle h, const dsb_symp_stage_args *a, int phase, const double *f0,
                                 const double *f1, void *stream)
{
    return symp_commit_events(h, a, phase, f0, f1, nullptr, stream);
}

int dsb_symp_stage_commit_events_located(dsb_symp_handle h, const dsb_symp_stage_args *a, const double *f0, const double *f1,
                                         const double *roots, void *stream)
{
    if (!roots) { set_error("dsb_symp_stage_commit_events_located: null roots"); return DSB_ERR_BAD_ARGUMENT; }
    return symp_commit_events(h, a, 1, f0, f1, roots, stream);
}

static int symp_commit_events(dsb_symp_handle h, const dsb_symp_stage_args *a, int phase, const double *f0, const double *f1,
                              const double *roots, void *stream)
{
    int rc = symp_check(h, a, "dsb_symp_stage_commit_events");
    if (rc != DSB_OK) return rc;
    if (a->n_events < 1 || a->n_events > DSB_MAX_STAGE_EVENTS || !a->ev_coef || !a->ev_flags || !a->ev_records || !a->ev_count ||
        !a->ev_state || !a->tf_per || !a->ev_scratch || a->ev_capacity < 1 || (phase != 0 && phase != 1) ||
        (phase == 1 && (!f0 || !f1))) {
        set_error("dsb_symp_stage_commit_events: needs 1..%d events with all ev_* buffers, tf_per, phase 0 or 1, and f0, f1 in "
                  "phase 1", DSB_MAX_STAGE_EVENTS);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (a->n == 0) return DSB_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (phase == 0) {
        symp_advance_kernel<<<(unsigned)((a->n + 127) / 128), 128, 0, st>>>(*a);
        DSB_CUDA_CHECK(cudaGetLastError());
        return DSB_OK;
    }
    // the event kernel of erk_stage.cu over a view of the symplectic buffers: dY = dS, accepted = steps, the
    // "accepted" bit of that kernel is this file's "stepped" bit
    dsb_erk_stage_args v;
    memset(&v, 0, sizeof(v));
    v.n = a->n; v.dim = a->dim; v.chunks = 1;
    v.y = a->y; v.t = a->t; v.dt = a->dt; v.dY = a->dS; v.flags = a->flags; v.trial = a->ev_scratch;
    v.accepted = a->steps; v.status = a->status; v.counters = a->counters; v.tf = a->tf;
    v.n_events = a->n_events; v.ev_capacity = a->ev_capacity; v.ev_coef = a->ev_coef; v.ev_flags = a->ev_flags;
    v.ev_records = a->ev_records; v.ev_count = a->ev_count; v.ev_state = a->ev_state; v.tf_per = a->tf_per;
    v.ev_coef_dy = a->ev_coef_dy;
    rc = launch_stage_events(v, f0, f1, roots, st);
    if (rc != DSB_OK) return rc;
    const int grid = grid_for((long long)a->dim * a->n, 256, h->sm_count);
    if (h->flags & DSB_FLAG_STRICT) symp_commit_kernel<true, SFLAG_STEPPED><<<grid, 256, 0, st>>>(*a);
    else symp_commit_kernel<false, SFLAG_STEPPED><<<grid, 256, 0, st>>>(*a);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // extern "C"
