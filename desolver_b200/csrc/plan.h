// plan.h -- host-side handle behind dsb_erk_handle, shared by abi.cu and the kernel translation units.
#pragma once
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
    double2 *d_atan = nullptr;           // device copy of the atan reduction table
    double *d_tab = nullptr;             // device copy of the flat tableau for the stage-level kernels
    // type-erased fused launcher (nullptr: no fused kernel for this rhs/method)
    int (*fused_launch)(const dsb_erk_plan *, const dsb_erk_run_args *, cudaStream_t) = nullptr;
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

template <class Rhs, int S, bool STRICT>
int fused_trampoline(const dsb_erk_plan *p, const dsb_erk_run_args *a, cudaStream_t stream)
{
    ErkKernelArgs<S> args;
    args.a = *a;
    fill_tableau<S>(p, args.tab);
    args.atan_table = p->d_atan;
    return launch_erk_fused<Rhs, S, STRICT>(args, p->fused_blocks, stream);
}

// Selects the instantiation for (stages, strict) of one right-hand side.  Returns false if none.
template <template <bool> class RhsT>
bool select_fused(dsb_erk_plan *p)
{
    const bool strict = p->flags & DSB_FLAG_STRICT;
    int per_sm = 0;
#define DSB_CASE(SS)                                                                          \
    case SS:                                                                                  \
        if (strict) { p->fused_launch = fused_trampoline<RhsT<true>, SS, true>;               \
                      occupancy_erk_fused<RhsT<true>, SS, true>(&per_sm); }                   \
        else        { p->fused_launch = fused_trampoline<RhsT<false>, SS, false>;             \
                      occupancy_erk_fused<RhsT<false>, SS, false>(&per_sm); }                 \
        break;
    switch (p->stages) {
        DSB_CASE(1) DSB_CASE(2) DSB_CASE(4) DSB_CASE(6) DSB_CASE(7) DSB_CASE(13) DSB_CASE(17)
    default: return false;
    }
#undef DSB_CASE
    if (per_sm < 1) per_sm = 1;
    p->fused_blocks = p->sm_count * per_sm;
    return true;
}

bool select_fused_lorenz(dsb_erk_plan *p);
bool select_fused_vdp(dsb_erk_plan *p);
bool select_fused_sho(dsb_erk_plan *p);

}  // namespace dsb
