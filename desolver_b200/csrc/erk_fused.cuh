// erk_fused.cuh -- K1: the fused, persistent, thread-per-trajectory explicit Runge-Kutta
// ensemble kernel (sm_100a).
//
// One launch integrates every trajectory of the shard from its current (t, y, dt) to tf.  Each
// lane owns ONE trajectory: state, all S stage derivatives, controller history and counters live
// in registers; nothing but the initial load and the final store touches HBM.  A lane whose
// trajectory finishes pulls the next one from a global work queue (warp-aggregated atomicAdd), so
// the 189..286-attempt spread between trajectories costs no lock-step
// idling and no host round trip.  One loop iteration = one step ATTEMPT of every live lane:
//     stage accumulation   y + h * sum_j a_sj k_j      (runge_kutta_methods.py:44-54)
//     propagated increment h * sum_j b_j k_j           (integrator_types.py:307)
//     embedded error       h * sum_j (b-bhat)_j k_j    (integrator_types.py:190,321)
//     weighted L2 norm + controller + accept/reject    (integrator_template.py:46-93)
//     masked commit, final-step clamp, termination     (differential_system.py:996-1002,1015-1018,1070-1071)
//     retry chain h <- min(new, h0), <= 64 retries     (integrator_types.py:201-224)
// Per-trajectory semantics are those of the reference run on that trajectory alone.
#pragma once

#include "common.cuh"
#include "fastmath.cuh"
#include "factor_table.h"
#include <math_constants.h>
#include <type_traits>
#include "tableau_static.cuh"
#include "rhs_builtin.cuh"

namespace dsb {

#ifndef DSB_ERK_THREADS
#define DSB_ERK_THREADS 128
#endif
#ifndef DSB_ERK_REFILL_BATCH
#define DSB_ERK_REFILL_BATCH 4
#endif
constexpr int ERK_THREADS = DSB_ERK_THREADS;
constexpr int REFILL_BATCH = DSB_ERK_REFILL_BATCH;   // lanes of a warp that must be waiting before the warp stores + refills

template <int S>
struct ErkKernelArgs {
    dsb_erk_run_args a;
    TableauData<S> tab;
    const double2 *atan_table;   // device copy of fm::build_atan_table
    const ft::Entry *factor_table;   // device copy of ft::build_table for this method's order (adaptive, fast flavour)
};

// compile-time loop: f(std::integral_constant<int, I>) for I in [I0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// ---- tableau policies -------------------------------------------------------------------
// StaticTab<T>: T is one of the generated structs of tableau_static.cuh; coefficients are
// constant expressions, zero entries are skipped at compile time (exactly the reference's
// "non-zero coefficients only" rule of runge_kutta_methods.py:45-48).
template <class T>
struct StaticTab {
    static constexpr int S = T::S;
    // Which terms exist is decided at compile time from the generated constexpr tableau; the VALUES are
    // read from the kernel parameter (dsb_erk_create verified both copies are bit-identical), because a
    // constant-bank operand costs no instruction while a 64-bit immediate needs two UMOVs.
    const TableauData<T::S> &d;
    template <int s, int j> static constexpr bool use_a() { return T::A[s][j] != 0.0; }
    template <int s, int j> __device__ __forceinline__ double a() const { return d.a[s][j]; }
    template <int s> __device__ __forceinline__ double c() const { return d.c[s]; }
    template <int j> static constexpr bool use_b() { return T::B[j] != 0.0; }
    template <int j> __device__ __forceinline__ double b() const { return d.b[j]; }
    template <int j> static constexpr bool use_db() { return (T::B[j] - T::BH[j]) != 0.0; }
    template <int j> __device__ __forceinline__ double db() const { return d.db[j]; }
    __device__ __forceinline__ bool adaptive() const { return T::ADAPTIVE; }
    __device__ __forceinline__ bool fsal() const { return T::FSAL; }
    __device__ __forceinline__ double order() const { return T::ORDER; }
    __device__ __forceinline__ int order_p() const { return T::ORDER_P; }
    __device__ explicit StaticTab(const TableauData<T::S> &dd) : d(dd) {}
};

// RuntimeTab<S>: any explicit tableau with S stages, coefficients read from the kernel
// parameter (constant bank).  The strictly lower triangle is treated as dense: a zero
// coefficient contributes k*0 = +0, which leaves every finite sum unchanged bit for bit.
template <int S_>
struct RuntimeTab {
    static constexpr int S = S_;
    const TableauData<S_> &d;
    template <int s, int j> static constexpr bool use_a() { return j < s; }
    template <int s, int j> __device__ __forceinline__ double a() const { return d.a[s][j]; }
    template <int s> __device__ __forceinline__ double c() const { return d.c[s]; }
    template <int j> static constexpr bool use_b() { return true; }
    template <int j> __device__ __forceinline__ double b() const { return d.b[j]; }
    template <int j> static constexpr bool use_db() { return true; }
    template <int j> __device__ __forceinline__ double db() const { return d.db[j]; }
    __device__ __forceinline__ bool adaptive() const { return d.adaptive != 0; }
    __device__ __forceinline__ bool fsal() const { return d.fsal != 0; }
    __device__ __forceinline__ double order() const { return d.order; }
    __device__ __forceinline__ int order_p() const { return d.order_p; }
    __device__ explicit RuntimeTab(const TableauData<S_> &dd) : d(dd) {}
};

// numpy's add.reduce over a contiguous axis of length S (see oracle/desolver_oracle.c):
// S < 8 left to right; else 8 accumulators, pairwise combine, sequential tail.  S <= 17 here.
template <int S, bool STRICT>
__device__ __forceinline__ double np_sum_terms(const double (&term)[S])
{
    using A = Ar<STRICT>;
    if constexpr (S < 8) {
        double s = term[0];
#pragma unroll
        for (int i = 1; i < S; ++i) s = A::add(s, term[i]);
        return s;
    } else {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = term[j];
        constexpr int full = S - (S % 8);
#pragma unroll
        for (int i = 8; i < full; i += 8)
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = A::add(r[j], term[i + j]);
        double res = A::add(A::add(A::add(r[0], r[1]), A::add(r[2], r[3])), A::add(A::add(r[4], r[5]), A::add(r[6], r[7])));
#pragma unroll
        for (int i = full; i < S; ++i) res = A::add(res, term[i]);
        return res;
    }
}

// pow-with-guard of the reference's 3-term branch: where(k > 0, k, 1)
__device__ __forceinline__ double guarded_pow(double base, double expo)
{
    double k = pow(base, expo);
    return k > 0.0 ? k : 1.0;
}

// corr contribution of one squared error norm x = ||diff/tol||^2:
//   where(eps > 0, eps ** (1/order), 1)  with eps = 1/sqrt(x)     (integrator_template.py:64,74)
template <bool STRICT>
__device__ __forceinline__ double corr_term(double x, double order, int order_p)
{
    if (STRICT) {
        double eps = 1.0 / sqrt(x);
        return eps > 0.0 ? pow(eps, 1.0 / order) : 1.0;
    }
    // 2^-120 <= x < 2^120 (|err/tol| in [7.5e-19, 1.2e18]) tested on the exponent bits: two integer instructions
    // instead of two DSETP plus their constants; zero, denormal, inf and nan all fall outside
    const unsigned hi = (unsigned)__double2hiint(x);
    if (hi - 0x38700000u < 0x0f000000u && order_p > 0) {
        if (order_p == 10) return fm::inv_root_narrow<10>(x);
        if (order_p == 16) return fm::inv_root_narrow<16>(x);
        return fm::inv_root_rt(x, order_p);
    }
    if (x > 1e-280 && x < 1e280 && order_p > 0) return fm::inv_root_rt(x, order_p);
    // eps = inf: the reference gets corr = inf and factor 1 + atan(inf) = 1 + pi/2; 1e300 gives the same double
    // (atan(0.8e300 - 1) rounds to pi/2) and keeps the table-driven atan in range without a guard
    if (x == 0.0) return 1e300;
    if (!(x < CUDART_INF)) return 1.0;                  // inf or nan: eps = 0 or nan -> 1
    return pow(x, -0.5 / order);
}

// resident blocks per SM: 6 (<=80 registers, no spill in the hot loop; measured best for Cash-Karp x dim 3 since the
// factor table shortened the controller: 3.68 ms against 3.79 ms at 7 blocks / 72 registers), 7 for dim 2 (Van der Pol 2^24: 40.7 ms against 43.9 ms at 8 and
// 42.1 ms at 6), 4 (<=128)
// while the stage derivatives fit, else 2 (<=255)
template <class Rhs, int S>
constexpr int erk_min_blocks()
{
#ifdef DSB_ERK_MIN_BLOCKS
    return DSB_ERK_MIN_BLOCKS;
#else
    return (S * Rhs::DIM <= 12) ? 7 : ((S * Rhs::DIM <= 18) ? 6 : ((S * Rhs::DIM <= 21) ? 4 : 2));
#endif
}

// Flow-control hooks of dsb_erk_run_host, kept out of line so that their registers do not enter the allocation of the
// hot loop.  flow_wait: spin (bounded, ~2 s) until `*avail >= want`; returns false after a time-out.
static __device__ __noinline__ bool flow_wait(const uint32_t *avail, unsigned want)
{
    unsigned have;
    for (long long spins = 0; spins <= 4000000; ++spins) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(avail) : "memory");
        if (have >= want) return true;
        __nanosleep(500);
    }
    return false;
}
// flow_done: results of one trajectory are stored -> bump its chunk's completion counter.  Results and counter both live
// in device memory: L2 is the coherence point of the SMs, the stream wait and the copy engine, so a device-scope fence
// orders them (a system-scope fence here cost 0.6 ms per 2^20 trajectories).
static __device__ __noinline__ void flow_done(uint32_t *counter)
{
    __threadfence();
    atomicAdd(counter, 1u);
}

// FLOW: the launch overlaps its own uploads / downloads (dsb_erk_run_host): it waits for `avail` before loading a
// trajectory and counts stored results in `done`.  A separate instantiation, because even untaken the hooks cost the
// device-resident kernel 3.6 % (3.53 -> 3.66 ms: different register allocation in the hot loop).
// EV: event detection after every accepted step (dsb_erk_run_args.n_events > 0), see the event block below.
// (EVK: 0 none, 1 events g(t, y), 2 events that may also depend on dState/dt -- a separate instantiation, because the
// extra live values cost the plain event kernel 15 %: 6.0 -> 6.9 ms)
// SCATTER: results are stored through the out_* pointers at index idx * out_index_mul + out_index_add instead of
// y / t / dt / accepted / rejected / status at idx: the sharded ensemble's gather fused into the result store (the
// destination is the root rank's buffer, mapped through CUDA IPC; remote stores travel over NVLink while the rest
// of the shard is still integrating).  Needs fresh != 0 and y_init (the launch then reads nothing it writes).
// DENSE: one node (t, y, f = rhs(t, y)) per accepted step goes to the paged node store (dsb_node_store, node_store.cu):
// dense output and the per-step history of single systems.  Its own instantiation: the node's right-hand side and the
// page cursor would otherwise live in the hot loop of the plain kernel.
// MULTI: a list of stop times (dsb_erk_run_args.stops) instead of one final time: every trajectory integrates to
// stops[0], its state is recorded, it continues to stops[1] with integrate()'s entry logic re-applied, ... -- the
// reference's `for t in t_eval: ode_system.integrate(t)` (differential_system.py:1245-1248) inside ONE launch, bit for bit.
template <class Rhs, class Tab, bool STRICT, bool FLOW = false, int EVK = 0, bool SCATTER = false, bool DENSE = false,
          bool MULTI = false>
#ifndef DSB_EV_MAX_BLOCKS
#define DSB_EV_MAX_BLOCKS 5
#endif
__global__ void __launch_bounds__(ERK_THREADS, ((EVK || DENSE) ? (erk_min_blocks<Rhs, Tab::S>() > DSB_EV_MAX_BLOCKS ? DSB_EV_MAX_BLOCKS
                                                                                                         : erk_min_blocks<Rhs, Tab::S>())
                                                   : erk_min_blocks<Rhs, Tab::S>()))
erk_fused_kernel(const __grid_constant__ ErkKernelArgs<Tab::S> P)
{
    using A = Ar<STRICT>;
    constexpr bool EV = EVK != 0;
    static_assert(!(DENSE && (FLOW || SCATTER || EVK != 0)), "the dense-output kernel is device-resident and event-free");
    static_assert(!(MULTI && (FLOW || SCATTER || EVK != 0 || DENSE)), "the multi-stop kernel is a plain device-resident run");
    constexpr bool LANE_TF = EV || MULTI;          // the final time is a per-lane value
    constexpr int S = Tab::S, D = Rhs::DIM, NP = Rhs::NPARAM;
    constexpr unsigned FULL = 0xffffffffu;
    const Tab tab(P.tab);
    const long long n = P.a.n;
    const double tf = P.a.tf, rtol = P.a.rtol, atol = P.a.atol;
    const int lane = threadIdx.x & 31;
    const int step_budget = P.a.max_steps < 0 ? 0x7fffffff : (P.a.max_steps > 0x7fffffff ? 0x7fffffff : (int)P.a.max_steps);

    __shared__ double2 s_atan[fm::ATAN_TABLE_SIZE];
    // double2 slots per shared-memory entry: 4 used + 1 pad, so that the 128-bit loads of 32 lanes reading 32
    // different entries spread over all bank groups (80-byte stride: 4 wavefronts instead of 16; measured 3.81 ->
    // 3.77 ms).  With 8 resident blocks per SM the padded table would not fit, so those kernels stay unpadded.
#ifdef DSB_FT_SLOTS
    constexpr int FT_SLOTS = DSB_FT_SLOTS;
#else
    constexpr int FT_SLOTS = (EV || erk_min_blocks<Rhs, Tab::S>() >= 8) ? 4 : 5;   // EV: its own arrays need the room
#endif
    __shared__ double2 s_ft[(STRICT ? 1 : ft::HOT_ENTRIES) * FT_SLOTS];   // hot octaves of the step-size factor table
    // per-thread statistics slots (no atomics): attempts, accepted steps, failed / unfinished trajectories
    __shared__ unsigned long long s_att[ERK_THREADS], s_acc[ERK_THREADS];
    __shared__ unsigned s_fail[ERK_THREADS], s_run[ERK_THREADS];
    s_att[threadIdx.x] = 0; s_acc[threadIdx.x] = 0; s_fail[threadIdx.x] = 0; s_run[threadIdx.x] = 0;
    __shared__ int s_gave_up[ERK_THREADS / 32];          // FLOW: this warp's wait for its initial states timed out
    if (FLOW && lane == 0) s_gave_up[threadIdx.x >> 5] = 0;
    // event detection state (EV only): coefficients of the hyperplane events, and per thread g at the step start
    // and the time of each event's last recorded occurrence
    constexpr int NE = DSB_MAX_EVENTS, EVW = Rhs::DIM + 2;
    __shared__ double s_evc[EV ? NE * EVW : 1];
    __shared__ double s_evd[EV ? NE * Rhs::DIM : 1];         // weights on dState/dt (requires_dstate events)
    __shared__ int s_evf[EV ? NE : 1];
    __shared__ double s_gprev[EV ? NE : 1][EV ? ERK_THREADS : 1];
    __shared__ double s_evlast[EV ? NE : 1][EV ? ERK_THREADS : 1];
    const int ne = EV ? (P.a.n_events < NE ? P.a.n_events : NE) : 0;
    const bool ev_dy = EVK == 2 && P.a.ev_coef_dy != nullptr;  // some event also depends on dState/dt
    if (EV) {
        for (int i = threadIdx.x; i < ne * Rhs::DIM; i += blockDim.x) s_evd[i] = ev_dy ? P.a.ev_coef_dy[i] : 0.0;
        for (int i = threadIdx.x; i < ne * EVW; i += blockDim.x) s_evc[i] = P.a.ev_coef[i];
        for (int i = threadIdx.x; i < ne; i += blockDim.x) s_evf[i] = P.a.ev_flags[i];
        __syncthreads();
    }
    if (!STRICT) {
        for (int i = threadIdx.x; i < fm::ATAN_TABLE_SIZE; i += blockDim.x) s_atan[i] = P.atan_table[i];
        if (tab.adaptive()) {
            const double2 *src = reinterpret_cast<const double2 *>(P.factor_table + ft::HOT_OFFSET);
            for (int i = threadIdx.x; i < ft::HOT_ENTRIES * 4; i += blockDim.x) s_ft[(i >> 2) * FT_SLOTS + (i & 3)] = src[i];
        }
        __syncthreads();
    }
    const long long ld = P.a.stride > 0 ? P.a.stride : n;          // leading dimension of y (and y_init)
    const double *y_src = P.a.y_init ? P.a.y_init : P.a.y;
    const bool fresh = P.a.fresh != 0;

    // ---- per-lane trajectory state (registers) ----
    long long idx = 0;
    bool live = false, exhausted = false, final_step = false;
    double y[D], prm[NP > 0 ? NP : 1], k[S][D], scale[D];
    double t = 0.0, dt = 0.0, h = 0.0, h0 = 0.0;
    double corr_first = 1.0, x_last = 1.0, x_last_last = 1.0;
    int trial = 0, acc = 0, rej = 0;
    int nk = 0;                                   // DENSE: nodes this trajectory has recorded so far (ordinal of the next)
    long long pg_base = -1;                       // DENSE: first entry of the page this warp is filling (-1: none)
    int pg_used = 0;                              //        entries of it already used (both warp-uniform)
#pragma unroll
    for (int d = 0; d < D; ++d) { y[d] = 0.0; scale[d] = 0.0; }
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int d = 0; d < D; ++d) k[s][d] = 0.0;

    auto store = [&](int st) {
        if constexpr (SCATTER) {
            const long long o = idx * P.a.out_index_mul + P.a.out_index_add;
#pragma unroll
            for (int d = 0; d < D; ++d) P.a.out_y[(long long)d * P.a.out_stride + o] = y[d];
            P.a.out_t[o] = t;
            P.a.out_dt[o] = dt;
            P.a.out_accepted[o] = acc;
            P.a.out_rejected[o] = rej;
            P.a.out_status[o] = st;
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) P.a.y[(long long)d * ld + idx] = y[d];
            P.a.t[idx] = t;
            P.a.dt[idx] = dt;
            if (P.a.accepted) P.a.accepted[idx] = fresh ? acc : P.a.accepted[idx] + acc;
            if (P.a.rejected) P.a.rejected[idx] = fresh ? rej : P.a.rejected[idx] + rej;
            if (P.a.status) P.a.status[idx] = st;
        }
        if constexpr (DENSE) P.a.store.node_count[idx] = nk;
        s_att[threadIdx.x] += (unsigned long long)(acc + rej);
        s_acc[threadIdx.x] += (unsigned long long)acc;
        s_fail[threadIdx.x] += st == (int)DSB_TRAJ_TOL_FAIL;
        s_run[threadIdx.x] += st == (int)DSB_TRAJ_RUNNING;
        // dsb_erk_run_host: a download stream waits for "every trajectory of this chunk is stored"
        if (FLOW && P.a.done) flow_done(P.a.done + (idx >> P.a.done_shift));
    };
    // per-warp totals -> counters[2..5] (one atomic per statistic per warp, at warp exit)
    auto flush_stats = [&]() {
        __syncwarp();
        unsigned long long a = s_att[threadIdx.x], c = s_acc[threadIdx.x];
        unsigned f = s_fail[threadIdx.x], r = s_run[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(FULL, a, o);
            c += __shfl_xor_sync(FULL, c, o);
        }
        f = __reduce_add_sync(FULL, f);
        r = __reduce_add_sync(FULL, r);
        if (lane == 0) {
            unsigned long long *ctr = (unsigned long long *)P.a.counters;
            if (a) atomicAdd(ctr + 2, a);
            if (c) atomicAdd(ctr + 3, c);
            if (f) atomicAdd(ctr + 4, (unsigned long long)f);
            if (r) atomicAdd(ctr + 5, (unsigned long long)r);
        }
    };
    // DENSE, warp-collective: the lanes with `want` append the node (idx, nk, t, y, f) to the warp's page of the node
    // store, packed into consecutive entries -- each field row of a 32-entry group is written as one or two contiguous
    // segments, whatever trajectories the lanes hold.  A new page costs the warp one atomicAdd per page_entries nodes.
    auto dense_append = [&](bool want, const double (&f)[D]) {
        const unsigned m = __ballot_sync(FULL, want);
        if (!m) return;
        const dsb_node_store &st = P.a.store;
        const int cnt = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
        const int room = pg_base >= 0 ? st.page_entries - pg_used : 0;
        long long e;
        if (cnt <= room) {
            e = pg_base + pg_used + rank;
            pg_used += cnt;
        } else {
            unsigned long long nb = 0;
            if (lane == 0) {
                if (pg_base >= 0) st.page_fill[pg_base / st.page_entries] = st.page_entries;   // full after this append
                nb = atomicAdd((unsigned long long *)st.cursor, (unsigned long long)st.page_entries);
            }
            nb = __shfl_sync(FULL, nb, 0);
            const bool ok = (long long)(nb + st.page_entries) <= st.capacity;
            e = rank < room ? pg_base + pg_used + rank : (ok ? (long long)nb + (rank - room) : -1);
            pg_base = ok ? (long long)nb : -1;
            pg_used = ok ? cnt - room : 0;
        }
        if (want) {
            if (e >= 0) {
                const long long g = e >> 5;
                double *rec = st.pool + g * ((2 + 2 * D) * 32) + (e & 31);
                *reinterpret_cast<int2 *>(rec) = make_int2((int)idx, nk);
                rec[32] = t;
#pragma unroll
                for (int d = 0; d < D; ++d) { rec[(2 + d) * 32] = y[d]; rec[(2 + D + d) * 32] = f[d]; }
            } else {
                atomicAdd((unsigned long long *)st.cursor + 1, 1ull);          // pool full: dropped, but counted
            }
            nk++;
        }
    };
    auto dense_flush = [&]() {                    // at warp exit: the page in progress keeps its fill level
        if (DENSE && lane == 0 && pg_base >= 0) P.a.store.page_fill[pg_base / P.a.store.page_entries] = pg_used;
    };

    // Start of a step: final-step clamp (differential_system.py:997-1002).  Called when a trajectory is loaded
    // and after every accepted step, so the loop top has no per-attempt bookkeeping.
    double tf_lane = tf;               // EV: a terminal event re-targets the trajectory to the event time
    bool final_leg = false;            // EV: re-integration to a terminal root in progress (events are off, :1054)
    int kstop = 0;                     // MULTI: index of the stop this trajectory is heading for
    auto start_step = [&]() {
        const double tfc = LANE_TF ? tf_lane : tf;
        final_step = fabs(dt + t) > fabs(tfc);
        h = final_step ? (tfc - t) : dt;
    };
    // MULTI: the trajectory stands at stops[kstop] (record = true: write its state there) or starts (kstop = -1): move to
    // the next stop with integrate()'s entry logic (differential_system.py:956-957, 971-974); false = no stop is left
    auto advance_stop = [&](bool record) -> bool {
        for (;;) {
            if (record) {
#pragma unroll
                for (int d = 0; d < D; ++d) P.a.stop_y[((long long)kstop * D + d) * P.a.stop_stride + idx] = y[d];
            }
            if (kstop + 1 >= P.a.n_stops) return false;
            kstop++;
            tf_lane = P.a.stops[kstop];
            const double span = tf_lane - t;
            record = true;
            if (fabs(span) < DSB_EPS4) continue;                         // integrate() returns at once: same state
            if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;
            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
            if (!((dt != 0.0) && (fabs(span) >= DSB_EPS32))) continue;   // the while condition (:996) fails: same state
            return true;
        }
    };
    // g_e(t, y) of the hyperplane events
    auto ev_g = [&](int e, double tt, const double (&yy)[D]) -> double {
        double g = s_evc[e * EVW + D] * tt - s_evc[e * EVW + D + 1];
#pragma unroll
        for (int d = 0; d < D; ++d) g = fma(s_evc[e * EVW + d], yy[d], g);
        return g;
    };
    // wdy_e . v: the part of a requires_dstate event that sees dState/dt (v = a right-hand side, or a state for the
    // P, Q terms of the interpolant's derivative)
    auto ev_d = [&](int e, const double (&v)[D]) -> double {
        double g = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) g = fma(s_evd[e * D + d], v[d], g);
        return g;
    };

    // Rare path, entered by the whole warp when ANY lane rejected, finished, failed or is idle:
    // retry bookkeeping, result stores, refill from the global queue.  Returns false when the warp is out of work.
    bool redo = false, done = false;
    double new_dt = 0.0, x = 0.0, corr = 1.0, sc_cur[D];
    auto event_path = [&]() -> bool {
        if (live && redo) {
            // integrator_types.py:201-224: retry with min(new, h0); history for the 2- and 3-term controller
            rej++;
            if (trial == 0) {
                h0 = h;
                // (the fast flavour forms the two-term product of the next attempt from x_last, see the controller)
                if (STRICT) corr_first = corr;
            }
            trial++;
            x_last_last = x_last;
            x_last = x;
#pragma unroll
            for (int d = 0; d < D; ++d) scale[d] = sc_cur[d];
            if (trial > DSB_MAX_RETRIES) {
                store((int)DSB_TRAJ_TOL_FAIL);                       // state frozen at the step start
                live = false;
                trial = 0;
            } else {
                h = h0 >= 0.0 ? fmin(new_dt, h0) : -fmin(fabs(new_dt), fabs(h0));
            }
        }
        if (!live && done) {                                             // finished earlier, result still in registers
            bool more = (dt != 0.0) && (fabs((LANE_TF ? tf_lane : tf) - t) >= DSB_EPS32);
            bool on = false;
            if constexpr (MULTI) on = !more && advance_stop(true);       // reached a stop: record it, head for the next
            if (on) {
                start_step();
                live = true;
                trial = 0;
            } else {
                store(more ? (int)DSB_TRAJ_RUNNING : ((EV && final_leg) ? (int)DSB_TRAJ_EVENT : (int)DSB_TRAJ_DONE));
            }
            done = false;
        }
        // refill idle lanes from the queue (warp-aggregated atomicAdd); a loaded trajectory that has nothing to
        // do (already at tf, failed earlier) is stored at once and the lane asks again
        for (;;) {
            unsigned need = __ballot_sync(FULL, !live);
            if (!need || exhausted) break;
            int cnt = __popc(need);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd((unsigned long long *)P.a.counters, (unsigned long long)cnt);
            base = __shfl_sync(FULL, base, 0);
            exhausted = (long long)(base + cnt) >= n;
            bool abandon = false;
            if (FLOW && P.a.avail && (long long)base < n) {
                // dsb_erk_run_host: the initial states arrive chunk by chunk while this kernel runs; wait (bounded:
                // ~2 s) until everything this warp just claimed is resident.  After a time-out the warp keeps
                // claiming and counts the abandoned trajectories as done, so that no download stream waits forever;
                // counters[6] tells the host.
                const unsigned want = (unsigned)((long long)(base + cnt) < n ? (long long)(base + cnt) : n);
                int bad = s_gave_up[threadIdx.x >> 5];
                if (lane == 0 && !bad && !flow_wait(P.a.avail, want)) {
                    bad = 1;
                    s_gave_up[threadIdx.x >> 5] = 1;
                    atomicExch((unsigned long long *)P.a.counters + 6, 1ull);
                }
                abandon = __shfl_sync(FULL, bad, 0) != 0;
                if (abandon) s_gave_up[threadIdx.x >> 5] = 1;
            }
            long long mine = (long long)base + __popc(need & ((1u << lane) - 1u));
            if (FLOW && abandon) {
                if (!live && mine < n && P.a.done) atomicAdd(P.a.done + (mine >> P.a.done_shift), 1u);
                continue;
            }
            bool want0 = false;                   // DENSE: this lane just loaded a trajectory without nodes
            double f0[DENSE ? D : 1];
            if (!live && mine < n) {
                idx = mine;
#pragma unroll
                for (int d = 0; d < D; ++d)   // FLOW: the copy engine writes this buffer while the kernel runs -> bypass L1
                    y[d] = FLOW ? __ldcg(y_src + (long long)d * ld + idx) : y_src[(long long)d * ld + idx];
                t = fresh ? P.a.t_init : P.a.t[idx];
                dt = fresh ? P.a.dt_init : P.a.dt[idx];
#pragma unroll
                for (int p = 0; p < NP; ++p) prm[p] = P.a.params[p][idx * P.a.param_stride[p]];
                Rhs::prepare(prm);
                acc = 0; rej = 0; trial = 0;
                if constexpr (DENSE) nk = fresh ? 0 : P.a.store.node_count[idx];
                int st0 = (P.a.status && !fresh) ? P.a.status[idx] : (int)DSB_TRAJ_DONE;
                bool go = st0 != (int)DSB_TRAJ_TOL_FAIL && st0 != (int)DSB_TRAJ_EVENT;
                if (EV) {
                    tf_lane = tf;
                    final_leg = false;
                    double fl0[D];
                    if (ev_dy) eval_rhs<Rhs>(t, y, prm, fl0);
                    for (int e = 0; e < ne; ++e) {
                        s_gprev[e][threadIdx.x] = ev_g(e, t, y) + (ev_dy ? ev_d(e, fl0) : 0.0);
                        s_evlast[e][threadIdx.x] = CUDART_NAN;
                    }
                    if (fresh) P.a.ev_count[idx] = 0;
                }
                if constexpr (MULTI) {
                    kstop = -1;
                    if (go) go = advance_stop(false);                          // entry logic towards stops[0] (or further)
                } else {
                    if (go && P.a.first_call) {
                        // integrate() entry: differential_system.py:956-957, 971-974
                        double span = tf - t;
                        if (fabs(span) < DSB_EPS4) go = false;
                        else {
                            if (dt != 0.0 && ((dt > 0.0) != (span > 0.0))) dt = -dt;
                            if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
                        }
                    }
                    go = go && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32);     // :996
                }
                if (go && P.a.max_steps == 0) go = false;
                if (go) {
                    live = true;
                    start_step();
                    if constexpr (DENSE) {
                        if (nk == 0) {            // node 0 = the initial condition
                            eval_rhs<Rhs>(t, y, prm, f0);
                            want0 = true;
                        }
                    }
                } else {
                    const bool frozen = st0 == (int)DSB_TRAJ_TOL_FAIL || st0 == (int)DSB_TRAJ_EVENT;
                    bool more = !MULTI && !frozen && (dt != 0.0) && (fabs(tf - t) >= DSB_EPS32);
                    store(frozen ? st0 : (more ? (int)DSB_TRAJ_RUNNING : (int)DSB_TRAJ_DONE));
                }
            }
            if constexpr (DENSE) dense_append(want0, f0);
        }
        if (!live) {
            // permanently idle lane (queue exhausted): park it on a benign state so that its (discarded)
            // arithmetic keeps taking the same branches as the live lanes of the warp
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = 1.0;
            t = 0.0; dt = 1e-3; h = 1e-3; final_step = false; trial = 0;
        }
        return __any_sync(FULL, live);
    };

    if (!event_path()) { dense_flush(); flush_stats(); return; }

    for (;;) {
        // ------------------------------------------------------------ stages
        double dlast[D];
        static_for<0, S>([&](auto s_) {
            constexpr int s = decltype(s_)::value;
            double ys[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double sum = 0.0;
                static_for<0, s>([&](auto j_) {
                    constexpr int j = decltype(j_)::value;
                    if constexpr (Tab::template use_a<s, j>()) sum = A::mad(k[j][d], tab.template a<s, j>(), sum);
                });
                if (STRICT) {
                    double inc = A::mul(h, sum);
                    if (s == S - 1) dlast[d] = inc;
                    ys[d] = A::add(y[d], inc);
                } else {
                    if (s == S - 1) dlast[d] = h * sum;
                    ys[d] = (s == 0) ? y[d] : fma(h, sum, y[d]);      // stage 0: y + h*0
                }
            }
            // stage time t + c_s h (runge_kutta_methods.py:50); dead code for an autonomous right-hand side
            double ts = t;
            if constexpr (Rhs::TIME) ts = STRICT ? A::add(t, A::mul(h, tab.template c<s>())) : fma(h, tab.template c<s>(), t);
            eval_rhs<Rhs>(ts, ys, prm, k[s]);
        });

        // ------------------------------------------------------------ increment + error estimate
        double dY[D], err[D], slope[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (STRICT) {
                double tb[S], te[S];
                static_for<0, S>([&](auto j_) {
                    constexpr int j = decltype(j_)::value;
                    tb[j] = 0.0; te[j] = 0.0;
                    if constexpr (Tab::template use_b<j>()) tb[j] = A::mul(k[j][d], tab.template b<j>());
                    if constexpr (Tab::template use_db<j>()) te[j] = A::mul(tab.template db<j>(), k[j][d]);
                });
                double sb = np_sum_terms<S, STRICT>(tb);
                dY[d] = tab.fsal() ? dlast[d] : A::mul(h, sb);
                err[d] = A::mul(h, np_sum_terms<S, STRICT>(te));
                slope[d] = dY[d] / h;                                  // dState / timestep
            } else {
                double sb = 0.0, se = 0.0;
                static_for<0, S>([&](auto j_) {
                    constexpr int j = decltype(j_)::value;
                    if constexpr (Tab::template use_b<j>()) sb = fma(k[j][d], tab.template b<j>(), sb);
                    if constexpr (Tab::template use_db<j>()) se = fma(k[j][d], tab.template db<j>(), se);
                });
                dY[d] = tab.fsal() ? dlast[d] : sb;                    // non-FSAL: the commit does y = fma(h, sb, y)
                err[d] = se;                                           // the factor h enters the norm once: x = h^2 * sum(..)
                slope[d] = sb;                                         // = dY/h to rounding
            }
        }

        // ------------------------------------------------------------ controller
        redo = false;
        new_dt = h;
        if (tab.adaptive()) {
            const bool any_retry = __any_sync(FULL, trial > 0);       // warp-uniform: rare paths are skipped
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double sc = STRICT ? np_maximum(fabs(y[d]), fabs(slope[d])) : fm::abs_max_setp(y[d], slope[d]);
                if (any_retry) {                                       // EMA on retries (:58-59); warp-uniform branch
                    __syncwarp();                                      // keeps it a branch (no if-conversion of the EMA)
                    if (trial > 0) sc = A::add(A::mul(DSB_KC(c08, 0.8), scale[d]), A::mul(DSB_KC(c02, 0.2), sc));
                }
                sc_cur[d] = sc;
                double tol = A::mad(rtol, sc, atol);                  // atol + rtol*sc
                double q = STRICT ? err[d] / tol : fm::quick_div(err[d], tol);
                x = (d == 0) ? __dmul_rn(q, q) : fma(q, q, x);        // OpenBLAS ddot: FMA chain
            }
            if (!STRICT) x *= h * h;                                   // fast flavour: err[d] above is the estimate / h
            // retry chains (:76-90): corr of this attempt times the history; lanes on a first attempt skip it
            auto retry_corr = [&](double c_cur) -> double {
                if (trial == 1) return A::mul(c_cur, corr_first);      // 2nd attempt: two-term product (:76-79)
                // 3rd+ attempt: exponents 0.55/-0.27/0.05 over order (integrator_template.py:80-90)
                const double lo = 1e-250, hi = 1e250;
                const bool tame = x > lo && x < hi && x_last > lo && x_last < hi && x_last_last > lo && x_last_last < hi;
                if (!STRICT && tame && tab.order_p() > 0 && (tab.order_p() % 2) == 0)
                    return fm::corr_three_term(x, x_last, x_last_last, 20 * tab.order_p());
                double k1 = guarded_pow(1.0 / sqrt(x), 0.55 / tab.order());
                double k2 = guarded_pow(1.0 / sqrt(x_last), -0.27 / tab.order());
                double k3 = guarded_pow(1.0 / sqrt(x_last_last), 0.05 / tab.order());
                return A::mul(A::mul(k1, k2), k3);
            };
            double f;
            if constexpr (STRICT) {
                corr = corr_term<STRICT>(x, tab.order(), tab.order_p());   // first attempt of a step (:73-75)
                if (any_retry) {                                           // retry chains only; warp-uniform branch
                    __syncwarp();
                    if (trial >= 1) corr = retry_corr(corr);
                }
                double u = A::sub(A::mul(DSB_KC(c08, 0.8), corr), 1.0);
                f = A::add(1.0, atan(u));
            } else {
                // First attempt of a step: f = F(x) = 1 + atan(0.8 x^(-1/(2 order)) - 1) from the factor table
                // (factor_table.h), a pure function of the bits of its argument.  Second attempt (:76-79):
                // corr = eps_cur^(1/order) * eps_last^(1/order) = (x * x_last)^(-1/(2 order)), i.e. the SAME function
                // of the product of the two squared norms.  Third and later attempts (three-term product with
                // non-integer exponents, :80-90) are rare and go through fastmath.cuh.  Idle lanes read entry 0.
                double xe = x;
                if (any_retry) {                                           // warp-uniform branch
                    __syncwarp();
                    if (trial == 1) xe = x * x_last;
                }
                const bool deep = trial >= 2;
                const unsigned xhi = (unsigned)__double2hiint(xe);
                const int kt = ft::index_of_hi(xhi);
                const unsigned kh = (live && !deep) ? (unsigned)(kt - ft::HOT_OFFSET) : 0u;
                const double dx = xe - __hiloint2double((int)ft::midpoint_hi(xhi), 0);     // exact
                double cf[ft::DEG + 1];
                if (__all_sync(FULL, kh < (unsigned)ft::HOT_ENTRIES)) {
                    const double2 *e = &s_ft[kh * FT_SLOTS];
#pragma unroll
                    for (int i = 0; i < (ft::DEG + 1) / 2; ++i) { double2 v = e[i]; cf[2 * i] = v.x; cf[2 * i + 1] = v.y; }
                    f = ft::horner(cf, dx);
                } else {
                    // some lane left the hot octaves (first steps, truncated last step): global copy of the table,
                    // same coefficients, same Horner chain -> same bits; beyond the table: fastmath.cuh
                    const bool in_tab = (unsigned)kt < (unsigned)ft::ENTRIES;
                    const double2 *e = reinterpret_cast<const double2 *>(P.factor_table + (in_tab ? kt : 0));
#pragma unroll
                    for (int i = 0; i < (ft::DEG + 1) / 2; ++i) { double2 v = __ldg(e + i); cf[2 * i] = v.x; cf[2 * i + 1] = v.y; }
                    f = ft::horner(cf, dx);
                    if (!in_tab && !deep && live) {
                        // xe outside [2^-100, 2^28), zero, inf or nan.  For trial == 1 the two-term product is formed
                        // from the factors so that an overflowing / underflowing xe cannot change the result class.
                        corr = corr_term<STRICT>(x, tab.order(), tab.order_p());
                        if (trial == 1) corr = corr * corr_term<STRICT>(x_last, tab.order(), tab.order_p());
                        f = fm::one_plus_atan(fma(DSB_KC(c08, 0.8), fmin(corr, 1e300), -1.0), s_atan);
                    }
                }
                if (any_retry) {                                           // retry chains only; warp-uniform branch
                    __syncwarp();
                    if (deep) {
                        corr = retry_corr(1.0);
                        f = fm::one_plus_atan(fma(DSB_KC(c08, 0.8), fmin(corr, 1e300), -1.0), s_atan);
                    }
                }
            }
            new_dt = A::mul(f, h);
            redo = f < DSB_KC(reject_below, DSB_REJECT_BELOW);
        }

        // ------------------------------------------------------------ accept: commit and start the next step
        unsigned ev_mask = 0;
        bool node_now = false;                                          // DENSE: this lane accepted a step in this iteration
        double fnode[DENSE ? D : 1];
        double yprev[D], tprev = t, dtprev = dt;                        // EV: the state before the commit (interpolant,
#pragma unroll                                                          // roll-back); dead right after the event block
        for (int d = 0; d < D; ++d) yprev[d] = y[d];
        if (live && !redo) {
#pragma unroll
            for (int d = 0; d < D; ++d)
                y[d] = (STRICT || tab.fsal()) ? A::add(y[d], dY[d]) : fma(h, dY[d], y[d]);   // :1015
            t = A::add(t, h);                                                                 // :1016
            acc++;
            trial = 0;
            if (!final_step) dt = new_dt;                                                     // :1070-1071
            if constexpr (DENSE) {
                eval_rhs<Rhs>(t, y, prm, fnode);                                              // the reference's final_rhs
                node_now = true;
            }
            if (EV && !final_leg) {
                // an event is searched for iff g changes sign over the step (utilities/optimizer.py:252)
                double fc[D];
                if (ev_dy) eval_rhs<Rhs>(t, y, prm, fc);                 // requires_dstate events see rhs at the step end
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    if (e < ne) {
                        const double g1 = ev_g(e, t, y) + (ev_dy ? ev_d(e, fc) : 0.0), g0 = s_gprev[e][threadIdx.x];
                        s_gprev[e][threadIdx.x] = g1;
                        const int dir = s_evf[e] & 3;
                        const bool rising = g0 < 0.0;
                        if (g0 * g1 < 0.0 && (dir == 0 || (dir == 1) == rising)) ev_mask |= 1u << e;
                    }
                }
            }
            start_step();
            done = !((dt != 0.0) && (fabs((LANE_TF ? tf_lane : tf) - t) >= DSB_KC(eps32, DSB_EPS32)))   // :996
                   || acc >= step_budget;
            live = !done;
        }
        if constexpr (DENSE) {
            // the node carries the state AFTER the commit, but `t`/`y` of a lane that just finished are still those values
            dense_append(node_now, fnode);
        }
        // ------------------------------------------------------------ events (EV): locate, record, terminal roll-back
        if (EV) {
            if (__any_sync(FULL, ev_mask != 0u) && ev_mask != 0u) {
                const int tid = threadIdx.x;
                // the accepted step [t0, t1] and its cubic Hermite interpolant (integrator_types.py:109-119,
                // utilities/interpolation.py:41-61): p0, p1 = states, m0 = the step's first stage, m1 = rhs(t1, y1)
                const double t0 = tprev, t1 = t, trange = t1 - t0;
                double y0[D], f1[D];
#pragma unroll
                for (int d = 0; d < D; ++d) y0[d] = yprev[d];
                eval_rhs<Rhs>(t1, y, prm, f1);
                double root[NE];
                int which[NE], cnt = 0;
                for (int e = 0; e < ne; ++e) {
                    if (!(ev_mask >> e & 1u)) continue;
                    // along the interpolant a linear g is itself the Hermite cubic of (g0, g1, dg0, dg1):
                    //   G(tau) = g0 + a tau + c2 tau^2 + c3 tau^3,  a = trange*dg0, b = trange*dg1
                    const double gy0 = ev_g(e, t0, y0), gy1 = ev_g(e, t1, y);
                    double dg0 = s_evc[e * EVW + D], dg1 = dg0;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        dg0 = fma(s_evc[e * EVW + d], k[0][d], dg0);
                        dg1 = fma(s_evc[e * EVW + d], f1[d], dg1);
                    }
                    double g0 = gy0, a = trange * dg0;
                    const double b = trange * dg1;
                    double c2 = 3.0 * (gy1 - gy0) - 2.0 * a - b;
                    const double c3 = 2.0 * (gy0 - gy1) + a + b;
                    double g1 = gy1;
                    if (ev_dy) {
                        // + wdy . grad(tau) = M0 + tau (-6 (P - Q)/trange - 4 M0 - 2 M1) + tau^2 (6 (P - Q)/trange + 3 M0 + 3 M1)
                        // (utilities/interpolation.py:63-78)
                        const double M0 = ev_d(e, k[0]), M1 = ev_d(e, f1), s6 = 6.0 * (ev_d(e, y0) - ev_d(e, y)) / trange;
                        g0 += M0;
                        g1 += M1;
                        a += -s6 - 4.0 * M0 - 2.0 * M1;
                        c2 += s6 + 3.0 * M0 + 3.0 * M1;
                    }
                    // bracketed Newton on [0, 1] (the reference runs Brent's method on the same function)
                    double lo = 0.0, hi = 1.0, tau = g0 / (g0 - g1);
                    for (int it = 0; it < 64; ++it) {
                        const double G = fma(fma(fma(c3, tau, c2), tau, a), tau, g0);
                        if (G == 0.0) break;
                        if ((G < 0.0) == (g0 < 0.0)) lo = tau; else hi = tau;
                        const double dG = fma(fma(3.0 * c3, tau, 2.0 * c2), tau, a);
                        const double step = G / dG;
                        // converged: the Newton correction is below the spacing of tau.  (Testing the bracket first
                        // would bisect a still-wide bracket whenever rounding puts the last iterate on its edge.)
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
                // in the order of time (:152-155); stop after the first terminal event (:157-162)
                for (int i = 1; i < cnt; ++i)
                    for (int j = i; j > 0 && (trange >= 0.0 ? root[j] < root[j - 1] : root[j] > root[j - 1]); --j) {
                        double tr = root[j]; root[j] = root[j - 1]; root[j - 1] = tr;
                        int tw = which[j]; which[j] = which[j - 1]; which[j - 1] = tw;
                    }
                bool term = false;
                double term_root = 0.0;
                for (int i = 0; i < cnt && !term; ++i) {
                    const int e = which[i];
                    const double rt = root[i], last = s_evlast[e][tid];
                    // a repeat of the same occurrence is dropped (:1043-1048: |t - t_last| <= epsilon^0.7)
                    if (!(last == last) || fabs(rt - last) > 2.9e-11) {
                        s_evlast[e][tid] = rt;
                        const int c = P.a.ev_count[idx];
                        if (c < P.a.ev_capacity) {
                            double *rec = P.a.ev_records + ((long long)idx * P.a.ev_capacity + c) * EVW;
                            const double ta = (rt - t0) / trange;                  // interpolation.py:38-39
                            rec[0] = rt;
                            rec[1 + D] = (double)e;
                            if (ta == 0.0 || ta == 1.0) {
#pragma unroll
                                for (int d = 0; d < D; ++d) rec[1 + d] = ta == 0.0 ? y0[d] : y[d];
                            } else {
                                const double t2 = __dmul_rn(ta, ta), t3 = __dmul_rn(t2, ta);
                                const double h00 = __dadd_rn(__dsub_rn(__dmul_rn(2.0, t3), __dmul_rn(3.0, t2)), 1.0);
                                const double h10 = __dadd_rn(__dsub_rn(t3, __dmul_rn(2.0, t2)), ta);
                                const double h01 = __dadd_rn(__dmul_rn(-2.0, t3), __dmul_rn(3.0, t2));
                                const double h11 = __dsub_rn(t3, t2);
#pragma unroll
                                for (int d = 0; d < D; ++d)
                                    rec[1 + d] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h00, y0[d]),
                                                                               __dmul_rn(__dmul_rn(h10, trange), k[0][d])),
                                                                     __dmul_rn(h01, y[d])),
                                                           __dmul_rn(__dmul_rn(h11, trange), f1[d]));
                            }
                        }
                        P.a.ev_count[idx] = c + 1;
                    }
                    if (s_evf[e] & 4) { term = true; term_root = rt; }
                }
                if (term) {
                    // terminal event: drop the step and integrate to the root like the nested
                    // self.integrate(roots[-1]) of differential_system.py:1051-1055 (counter -= 1; entry clamp
                    // dt = |t_root - t| / 2 of :973-974; final-step clamp; no event detection on that leg)
#pragma unroll
                    for (int d = 0; d < D; ++d) y[d] = y0[d];
                    t = t0;
                    dt = dtprev;
                    acc--;
                    tf_lane = term_root;
                    final_leg = true;
                    const double span = tf_lane - t;
                    if (fabs(span) < DSB_EPS4) {                         // integrate() returns at once (:956-957)
                        done = true;
                    } else {
                        if (fabs(dt) > fabs(span)) dt = copysign(fabs(span) * 0.5, span);
                        start_step();
                        done = !((dt != 0.0) && (fabs(tf_lane - t) >= DSB_EPS32)) || acc >= step_budget;
                    }
                    live = !done;
                }
            }
        }
        // ------------------------------------------------------------ rare events: reject, finish, refill
        // A finished lane keeps its result in registers (its further arithmetic is discarded) until REFILL_BATCH
        // lanes of the warp need service, a lane rejected, or the queue is empty: the store + refill pass costs the
        // whole warp ~250 instructions, and trajectories that start together take their first-step retry chain
        // (dt0 is usually too large: two rejections, i.e. the expensive three-term branch) in the same iterations.
        const unsigned svc = __ballot_sync(FULL, !live && (done || !exhausted));
        if (__any_sync(FULL, live && redo) || __popc(svc) >= REFILL_BATCH || (exhausted && svc != 0u)) {
            if (!event_path()) break;
        }
    }
    dense_flush();
    flush_stats();
}

// Host-side launcher, one per instantiation.
template <class Rhs, class Tab, bool STRICT, bool FLOW = false, int EVK = 0, bool SCATTER = false, bool DENSE = false,
          bool MULTI = false>
int launch_erk_fused(const ErkKernelArgs<Tab::S> &args, int blocks, cudaStream_t stream)
{
    erk_fused_kernel<Rhs, Tab, STRICT, FLOW, EVK, SCATTER, DENSE, MULTI><<<blocks, ERK_THREADS, 0, stream>>>(args);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

template <class Rhs, class Tab, bool STRICT, bool FLOW = false, int EVK = 0, bool SCATTER = false, bool DENSE = false,
          bool MULTI = false>
int occupancy_erk_fused(int *blocks_per_sm)
{
    DSB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        blocks_per_sm, erk_fused_kernel<Rhs, Tab, STRICT, FLOW, EVK, SCATTER, DENSE, MULTI>, ERK_THREADS, 0));
    return DSB_OK;
}

}  // namespace dsb
