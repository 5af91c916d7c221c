// plan.h -- host-side handle behind dsb_erk_handle, shared by abi.cu and the kernel translation units.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "erk_fused.cuh"

struct dsb_erk_plan {
    int device = 0, rhs_id = 0, dim = 0, stages = 0, final_rows = 0;
    double order = 0.0;
    unsigned flags = 0;
    std::vector<double> inter, fin;      // host copies of the class tableaux
    int sm_count = 0;
    int fused_blocks = 0;                // persistent grid size (SMs x resident blocks)
    bool fused_static = false;           // compile-time tableau variant selected
    double2 *d_atan = nullptr;           // device copy of the atan reduction table
    double *d_tab = nullptr;             // device copy of the flat tableau for the stage-level kernels
    cudaEvent_t ev_start = nullptr, ev_done = nullptr;   // dsb_erk_run_host: fork / join of the copy streams
    dsb::ft::Entry *d_factor = nullptr;  // device copy of the step-size factor table (factor_table.h), adaptive only
    // type-erased fused launcher (nullptr: no fused kernel for this rhs/method)
    int (*fused_launch)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    // same kernel with the upload / download flow control of dsb_erk_run_host compiled in (fast flavour only)
    int (*fused_launch_flow)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    // same kernel with event detection compiled in (<= 7 stages); its own persistent grid size
    int (*fused_launch_events)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    int (*fused_launch_events_dy)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;   // + dState/dt events
    int events_blocks = 0, events_dy_blocks = 0;
    // same kernel storing its results through the out_* scatter pointers (fast flavour only)
    int (*fused_launch_scatter)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    // same kernel writing one node per accepted step to the paged node store (dense output / step history)
    int (*fused_launch_dense)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    int dense_blocks = 0;
    // same kernel heading for a list of stop times (solve_ivp's t_eval in one launch); <= 7 stages
    int (*fused_launch_multi)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
    int multi_blocks = 0;
};

namespace dsb {

template <int S>
void fill_tableau(const dsb_erk_plan *p, TableauData<S> &t)
{
    const int W = S + 1;
    t.nz_b = t.nz_db = 0;
    for (int s = 0; s < S; ++s) {
        t.c[s] = p->inter[s * W];
        t.nz_a[s] = 0;
        for (int j = 0; j < S; ++j) {
            t.a[s][j] = p->inter[s * W + 1 + j];
            if (t.a[s][j] != 0.0) t.nz_a[s] |= 1u << j;
        }
        t.b[s] = p->fin[1 + s];
        // (b - bhat) formed in fp64 at run time, like integrator_types.py:321
        t.db[s] = p->final_rows == 2 ? p->fin[1 + s] - p->fin[W + 1 + s] : 0.0;
        if (t.b[s] != 0.0) t.nz_b |= 1u << s;
        if (t.db[s] != 0.0) t.nz_db |= 1u << s;
    }
    t.adaptive = p->final_rows == 2;
    t.fsal = 1;
    for (int j = 0; j < S; ++j)
        if (t.a[S - 1][j] != t.b[j]) t.fsal = 0;
    t.order = p->order;
    double p2 = 2.0 * p->order;
    t.order_p = (p2 == (double)(int)p2 && p2 >= 1.0 && p2 <= 64.0) ? (int)p2 : 0;
}

// Persistent grid for an ensemble of n trajectories: `full` = SMs x resident blocks is right when the ensemble is many
// waves deep (the 2^20-trajectory benchmark: 9.2 waves).  A SMALL ensemble finishes sooner on fewer resident blocks:
// the run is (waves + ~1/2) trajectory latencies long -- never less than ~1.25 -- and a trajectory's latency grows with the
// warps that share a scheduler.  Measured for Cash-Karp x Lorenz (tools/grid_sweep.py, latency of one trajectory relative
// to one block per SM): 1.00, 1.24, 1.54, 1.89, 2.28, 2.71 for 1..6 blocks, i.e. about 1 + 0.34 (w - 1).  Ensembles of at
// least 3 waves keep the full grid; below that the block count per SM minimises the model
// (2^15 trajectories: 0.34 -> 0.23 ms, 2^16: 0.46 -> 0.34 ms, 2^17: 0.58 -> 0.57 ms).  DSB_ERK_BLOCKS_PER_SM overrides.
inline int grid_for_ensemble(const dsb_erk_plan *p, int full, long long n)
{
    const int max_per_sm = full / (p->sm_count > 0 ? p->sm_count : 1);
    if (max_per_sm <= 1) return full;
    if (const char *env = getenv("DSB_ERK_BLOCKS_PER_SM")) {
        int w = atoi(env);
        if (w >= 1 && w <= max_per_sm) return p->sm_count * w;
    }
    const double lanes_per_block_row = (double)p->sm_count * ERK_THREADS;
    if ((double)n >= 3.0 * lanes_per_block_row * max_per_sm) return full;
    double best = 1e300;
    int best_w = max_per_sm;
    for (int w = max_per_sm; w >= 1; --w) {
        const double waves = (double)n / (lanes_per_block_row * w);
        const double latency = 1.0 + 0.34 * (w - 1);
        const double rounds = waves + 0.55 > 1.25 ? waves + 0.55 : 1.25;
        const double cost = rounds * latency;
        if (cost < best * 0.995) { best = cost; best_w = w; }          // ties go to the larger grid
    }
    return p->sm_count * best_w;
}

template <class Rhs, class Tab, bool STRICT, bool FLOW = false, int EVK = 0, bool SCATTER = false, bool DENSE = false,
          bool MULTI = false>
int fused_trampoline(const dsb_erk_plan *p, const dsb_erk_run_args *a, cudaStream_t stream)
{
    ErkKernelArgs<Tab::S> args;
    args.a = *a;
    fill_tableau<Tab::S>(p, args.tab);
    args.atan_table = p->d_atan;
    args.factor_table = p->d_factor;
    int blocks = MULTI ? p->multi_blocks : DENSE ? p->dense_blocks : (EVK == 2 ? p->events_dy_blocks : (EVK == 1 ? p->events_blocks : p->fused_blocks));
    blocks = grid_for_ensemble(p, blocks, a->n);
    return launch_erk_fused<Rhs, Tab, STRICT, FLOW, EVK, SCATTER, DENSE, MULTI>(args, blocks, stream);
}

// true iff the caller's tableau equals the compile-time one bit for bit
template <class T>
bool matches_static(const dsb_erk_plan *p)
{
    constexpr int S = T::S;
    if (p->stages != S || (p->final_rows == 2) != T::ADAPTIVE || p->order != T::ORDER) return false;
    for (int s = 0; s < S; ++s) {
        if (memcmp(&p->inter[s * (S + 1)], &T::C[s], 8) != 0) return false;
        for (int j = 0; j < S; ++j)
            if (memcmp(&p->inter[s * (S + 1) + 1 + j], &T::A[s][j], 8) != 0) return false;
        if (memcmp(&p->fin[1 + s], &T::B[s], 8) != 0) return false;
        if (T::ADAPTIVE && memcmp(&p->fin[S + 1 + 1 + s], &T::BH[s], 8) != 0) return false;
    }
    return true;
}

template <class Rhs, class Tab, bool STRICT>
void bind_fused(dsb_erk_plan *p, int *per_sm)
{
    p->fused_launch = fused_trampoline<Rhs, Tab, STRICT>;
    if constexpr (!STRICT && Tab::S <= 7) p->fused_launch_flow = fused_trampoline<Rhs, Tab, STRICT, true>;
    if constexpr (!STRICT && Tab::S <= 7) p->fused_launch_scatter = fused_trampoline<Rhs, Tab, STRICT, false, 0, true>;
    if constexpr (Tab::S <= 7) {
        p->fused_launch_events = fused_trampoline<Rhs, Tab, STRICT, false, 1>;
        p->fused_launch_events_dy = fused_trampoline<Rhs, Tab, STRICT, false, 2>;
        int ev_per_sm = 0;
        occupancy_erk_fused<Rhs, Tab, STRICT, false, 1>(&ev_per_sm);
        p->events_blocks = p->sm_count * (ev_per_sm < 1 ? 1 : ev_per_sm);
        occupancy_erk_fused<Rhs, Tab, STRICT, false, 2>(&ev_per_sm);
        p->events_dy_blocks = p->sm_count * (ev_per_sm < 1 ? 1 : ev_per_sm);
    }
    {
        p->fused_launch_dense = fused_trampoline<Rhs, Tab, STRICT, false, 0, false, true>;
        int d_per_sm = 0;
        occupancy_erk_fused<Rhs, Tab, STRICT, false, 0, false, true>(&d_per_sm);
        p->dense_blocks = p->sm_count * (d_per_sm < 1 ? 1 : d_per_sm);
    }
    if constexpr (Tab::S <= 7) {
        p->fused_launch_multi = fused_trampoline<Rhs, Tab, STRICT, false, 0, false, false, true>;
        int m_per_sm = 0;
        occupancy_erk_fused<Rhs, Tab, STRICT, false, 0, false, false, true>(&m_per_sm);
        p->multi_blocks = p->sm_count * (m_per_sm < 1 ? 1 : m_per_sm);
    }
    occupancy_erk_fused<Rhs, Tab, STRICT>(per_sm);
}

// Selects the instantiation for (tableau, strict) of one right-hand side.  The flagship
// tableaux get compile-time coefficients; anything else with a supported stage count runs the
// runtime-coefficient kernel.  Returns false if there is none.
template <template <bool> class RhsT>
bool select_fused(dsb_erk_plan *p)
{
    const bool strict = p->flags & DSB_FLAG_STRICT;
    int per_sm = 0;
    p->fused_static = true;
#ifdef DSB_SELECT_FAST_ONLY
    // plug-ins built by rhs_plugins.register_device_rhs: the fast flavour only (half the instantiations; a STRICT system
    // of such a right-hand side runs on the stage-level kernels)
    if (strict) return false;
#define DSB_BIND(TAB) do { bind_fused<RhsT<false>, TAB, false>(p, &per_sm); } while (0)
#else
#define DSB_BIND(TAB)                                                                         \
    do { if (strict) bind_fused<RhsT<true>, TAB, true>(p, &per_sm);                           \
         else        bind_fused<RhsT<false>, TAB, false>(p, &per_sm); } while (0)
#endif
    if (matches_static<CashKarp45>(p)) DSB_BIND(StaticTab<CashKarp45>);
    else if (matches_static<DormandPrince87>(p)) DSB_BIND(StaticTab<DormandPrince87>);
    else if (matches_static<DormandPrince54>(p)) DSB_BIND(StaticTab<DormandPrince54>);
    else {
        p->fused_static = false;
        switch (p->stages) {
        case 1: DSB_BIND(RuntimeTab<1>); break;
        case 2: DSB_BIND(RuntimeTab<2>); break;
        case 4: DSB_BIND(RuntimeTab<4>); break;
        case 6: DSB_BIND(RuntimeTab<6>); break;
        case 7: DSB_BIND(RuntimeTab<7>); break;
        case 13: DSB_BIND(RuntimeTab<13>); break;
        case 17: DSB_BIND(RuntimeTab<17>); break;
        default: return false;
        }
    }
#undef DSB_BIND
    if (per_sm < 1) per_sm = 1;
    p->fused_blocks = p->sm_count * per_sm;
    return true;
}

bool select_fused_lorenz(dsb_erk_plan *p);
bool select_fused_vdp(dsb_erk_plan *p);
bool select_fused_sho(dsb_erk_plan *p);
bool select_fused_forced(dsb_erk_plan *p);
bool select_fused_duffing(dsb_erk_plan *p);

}  // namespace dsb
