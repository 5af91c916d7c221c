// Instantiations of the fused persistent ERK kernel (erk_fused.cuh) for the Duffing oscillator.
#include "plan.h"
#include "rhs_builtin.cuh"

namespace dsb {
bool select_fused_duffing(dsb_erk_plan *p) { return select_fused<DuffingRhs>(p); }
}  // namespace dsb
