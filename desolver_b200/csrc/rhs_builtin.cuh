// rhs_builtin.cuh -- fused device right-hand sides ("RHS plugins") for the thread-per-trajectory
// kernels.  Each mirrors, operation for operation, the tagged Python callable of the same name in
// desolver_b200/rhs_plugins.py (which is what the reference integrates to produce the golden
// fixtures).  STRICT keeps one rounding per operator; the fast flavour lets nvcc contract to DFMA.
#pragma once
#include "common.cuh"

namespace dsb {

template <bool STRICT>
struct LorenzRhs {
    static constexpr int ID = DSB_RHS_LORENZ, DIM = 3, NPARAM = 3;   // sigma, rho, beta
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&)[NPARAM]) {}
    static __device__ __forceinline__ void eval(const double (&y)[DIM], const double (&p)[NPARAM], double (&dy)[DIM])
    {
        dy[0] = A::mul(p[0], A::sub(y[1], y[0]));
        dy[1] = A::sub(A::mul(y[0], A::sub(p[1], y[2])), y[1]);
        dy[2] = A::sub(A::mul(y[0], y[1]), A::mul(p[2], y[2]));
    }
};

template <bool STRICT>
struct VdpRhs {
    static constexpr int ID = DSB_RHS_VDP, DIM = 2, NPARAM = 1;      // mu
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&)[NPARAM]) {}
    static __device__ __forceinline__ void eval(const double (&y)[DIM], const double (&p)[NPARAM], double (&dy)[DIM])
    {
        dy[0] = y[1];
        dy[1] = A::sub(A::mul(A::mul(p[0], A::sub(1.0, A::mul(y[0], y[0]))), y[1]), y[0]);
    }
};

template <bool STRICT>
struct ShoRhs {
    static constexpr int ID = DSB_RHS_SHO, DIM = 2, NPARAM = 2;      // k, m
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&p)[NPARAM]) { p[0] = -(p[0] / p[1]); }
    static __device__ __forceinline__ void eval(const double (&y)[DIM], const double (&p)[NPARAM], double (&dy)[DIM])
    {
        dy[0] = y[1];
        dy[1] = A::mul(p[0], y[0]);
    }
};

}  // namespace dsb
