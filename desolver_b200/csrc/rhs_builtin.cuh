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
    static constexpr bool TIME = false;
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
    static constexpr bool TIME = false;
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
    static constexpr bool TIME = false;
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&p)[NPARAM]) { p[0] = -(p[0] / p[1]); }
    static __device__ __forceinline__ void eval(const double (&y)[DIM], const double (&p)[NPARAM], double (&dy)[DIM])
    {
        dy[0] = y[1];
        dy[1] = A::mul(p[0], y[0]);
    }
};

// Time-dependent right-hand sides declare TIME = true and take the stage time t + c_s h
// (integrators/components/runge_kutta_methods.py:49-50) as the first argument of eval.

// y'' + y = t as a first-order system: the reference's own test fixture
// (/root/reference/desolver/tests/common.py:27-72; rhs_plugins.forced).
template <bool STRICT>
struct ForcedRhs {
    static constexpr int ID = DSB_RHS_FORCED, DIM = 2, NPARAM = 0;
    static constexpr bool TIME = true;
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&)[1]) {}
    static __device__ __forceinline__ void eval(double t, const double (&y)[DIM], const double (&)[1], double (&dy)[DIM])
    {
        dy[0] = y[1];
        dy[1] = A::add(-y[0], t);
    }
};

// Duffing oscillator x'' + delta x' + alpha x + beta x^3 = gamma cos(omega t)  (the system of the reference's
// docs/examples notebooks; rhs_plugins.duffing).  params: delta, alpha, beta, gamma, omega.
template <bool STRICT>
struct DuffingRhs {
    static constexpr int ID = DSB_RHS_DUFFING, DIM = 2, NPARAM = 5;
    static constexpr bool TIME = true;
    using A = Ar<STRICT>;
    static __device__ __forceinline__ void prepare(double (&)[NPARAM]) {}
    static __device__ __forceinline__ void eval(double t, const double (&y)[DIM], const double (&p)[NPARAM], double (&dy)[DIM])
    {
        // gamma*cos(omega*t) - delta*v - alpha*x - beta*x*x*x, left to right like the Python callable
        dy[0] = y[1];
        double f = A::mul(p[3], cos(A::mul(p[4], t)));
        f = A::sub(f, A::mul(p[0], y[1]));
        f = A::sub(f, A::mul(p[1], y[0]));
        dy[1] = A::sub(f, A::mul(A::mul(A::mul(p[2], y[0]), y[0]), y[0]));
    }
};

// uniform call site: autonomous right-hand sides ignore the time
template <class Rhs, int NPARR>
__device__ __forceinline__ void eval_rhs(double t, const double (&y)[Rhs::DIM], const double (&p)[NPARR], double (&dy)[Rhs::DIM])
{
    if constexpr (Rhs::TIME) Rhs::eval(t, y, p, dy);
    else Rhs::eval(y, p, dy);
}

}  // namespace dsb
