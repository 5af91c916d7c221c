// Instantiations of the fused persistent ERK kernel (erk_fused.cuh) for the lorenz right-hand side.
#include "plan.h"
#include "rhs_builtin.cuh"

namespace dsb {
bool select_fused_lorenz(dsb_erk_plan *p) { return select_fused<LorenzRhs>(p); }
}  // namespace dsb
