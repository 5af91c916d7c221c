// Instantiations of the fused persistent ERK kernel (erk_fused.cuh) for the vdp right-hand side.
#include "plan.h"
#include "rhs_builtin.cuh"

namespace dsb {
bool select_fused_vdp(dsb_erk_plan *p) { return select_fused<VdpRhs>(p); }
}  // namespace dsb
