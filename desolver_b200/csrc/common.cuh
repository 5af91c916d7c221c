// common.cuh -- shared definitions for the sm_100a kernels of libdesolver_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/desolver_b200.h"

namespace dsb {

// D.epsilon = 4 eps, D.tol_epsilon = 32 eps  (reference backend/autoray_backend.py:8-31)
#define DSB_EPS4 (4.0 * 2.220446049250313e-16)
#define DSB_EPS32 (32.0 * 2.220446049250313e-16)
// 0.9**2 in double; a step is rejected iff factor < this (integrator_template.py:93)
#define DSB_REJECT_BELOW (0.9 * 0.9)
#define DSB_MAX_RETRIES 64  // num_step_retries, integrator_types.py:148

void set_error(const char *fmt, ...);

#define DSB_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess) {                                                       \
            dsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),     \
                           __FILE__, __LINE__);                                           \
            return DSB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// Makes `device` current for the duration of an entry point and restores the caller's device afterwards: handles are
// cached per device on the Python side, and a launch on a stream of device 1 while device 0 is current fails with
// cudaErrorInvalidResourceHandle.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// Arithmetic policy.  STRICT: every + and * rounds separately (numpy's separate ufunc calls;
// nvcc never contracts the __d*_rn intrinsics).  Fast: plain operators, which nvcc contracts
// into DFMA (-fmad=true), halving the FP64 issue count.
template <bool STRICT>
struct Ar;

template <>
struct Ar<true> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    // a*b + c, two roundings
    static __device__ __forceinline__ double mad(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
};

template <>
struct Ar<false> {
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double mad(double a, double b, double c) { return fma(a, b, c); }
};

// np.maximum semantics (NaN-propagating); fmax would drop the NaN.
__device__ __forceinline__ double np_maximum(double a, double b)
{
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}

// Tableau as passed to kernels (by value, lives in the constant bank).  Layout follows the
// reference class data: inter[s*(S+1)] = c_s, inter[s*(S+1)+1+j] = a_sj; b = tableau_final[0,1:],
// db = tableau_final[0,1:] - tableau_final[1,1:] formed in fp64 on the host exactly like
// integrator_types.py:321 does at run time.
template <int S>
struct TableauData {
    double c[S];
    double a[S][S];
    double b[S];
    double db[S];
    unsigned nz_a[S];   // bit j set <=> a[s][j] != 0
    unsigned nz_b, nz_db;
    int adaptive;       // tableau_final has 2 rows
    int fsal;           // a[S-1][:] == b[:]  (integrator_types.py:134)
    int order_p;        // 2*order when that is an integer (fast inverse-root path), else 0
    double order;
};

}  // namespace dsb
