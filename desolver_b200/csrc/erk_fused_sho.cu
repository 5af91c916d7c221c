// Instantiations of the fused persistent ERK kernel (erk_fused.cuh) for the sho right-hand side.
#include "plan.h"
#include "rhs_builtin.cuh"

namespace dsb {
bool select_fused_sho(dsb_erk_plan *p) { return select_fused<ShoRhs>(p); }
}  // namespace dsb
