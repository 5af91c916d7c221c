// Instantiations of the fused persistent ERK kernel (erk_fused.cuh) for the time-dependent forced oscillator.
#include "plan.h"
#include "rhs_builtin.cuh"

namespace dsb {
bool select_fused_forced(dsb_erk_plan *p) { return select_fused<ForcedRhs>(p); }
}  // namespace dsb
