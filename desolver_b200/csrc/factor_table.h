// factor_table.h -- the step-size factor of the FIRST attempt of a step as ONE table-driven polynomial.
//
// On the first attempt of a step the reference controller (integrators/integrator_template.py:62-64,73-75,91)
// computes, from the squared weighted error norm x = ||diff/tol||_2^2,
//     eps  = 1/sqrt(x)                     :64
//     corr = eps ** (1/order)              :74
//     f    = 1 + arctan(0.8*corr - 1)      :91      (new_dt = f*h; the attempt is rejected iff f < 0.81, :93)
// i.e. a smooth scalar function F(x) = 1 + atan(0.8 * x^(-1/(2 order)) - 1).  Evaluating it as sqrt / reciprocal /
// pow / atan costs ~25 FP64 + ~34 other instructions per attempt even with the replacements of fastmath.cuh; the
// FP64 pipe and the issue slots are what bounds the fused kernel.  Here F is tabulated: each octave of
// x is cut into 2^M_BITS intervals by the leading mantissa bits, and on interval k
//     F(x_k + d) = c0 + d (c1 + d (c2 + ... + d c7)),     x_k = interval midpoint (exact double), d = x - x_k (exact),
// is the degree-7 Taylor polynomial about the midpoint (|d/x_k| <= 2^-6; the series in d/x_k has radius of
// convergence ~1, so the truncation error is < 1e-16 absolute: tests/fastmath_host.cpp measures <= 1 ulp against
// libm).  Cost on the device: 1 DADD + 7 DFMA + ~10 integer/load instructions.  The index comes from the bits of x,
// so the result is a pure function of x: bitwise deterministic, independent of which path loaded the entry.
//
// The table is built on the host in long double when a handle is created (dsb_erk_create), for that method's
// `order` and the reference's safety factor 0.8.  The kernel keeps the octaves [2^HOT_LO, 2^HOT_HI) in shared
// memory (99.4 % of the attempts of the Lorenz / Van der Pol ensembles) and reads the rest from global memory.
// Outside [2^E_LO, 2^E_HI), for x = 0 / inf / nan, and on retries (2- and 3-term products, :76-90) the
// controller falls back to fastmath.cuh.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace dsb {
namespace ft {

constexpr int M_BITS = 5;                      // 32 intervals per octave
constexpr int DEG = 7;
constexpr int E_LO = -100, E_HI = 28;          // the table covers 2^E_LO <= x < 2^E_HI
#ifndef DSB_FT_HOT_LO
#define DSB_FT_HOT_LO -8
#endif
constexpr int HOT_LO = DSB_FT_HOT_LO, HOT_HI = 2;   // octaves kept in shared memory
constexpr int ENTRIES = (E_HI - E_LO) << M_BITS;
constexpr int HOT_ENTRIES = (HOT_HI - HOT_LO) << M_BITS;
constexpr int HOT_OFFSET = (HOT_LO - E_LO) << M_BITS;
// index of x = (hi32(x) >> (20 - M_BITS)) - INDEX_BASE, valid iff (unsigned) index < ENTRIES
constexpr int INDEX_SHIFT = 20 - M_BITS;
constexpr int INDEX_BASE = (1023 + E_LO) << M_BITS;

struct alignas(64) Entry {
    double c[DEG + 1];
};
static_assert(sizeof(Entry) == 64, "one entry = four 128-bit loads");

// ---- host side: table construction (long double power-series arithmetic) ----------------------------------
inline void build_entry(int k, double order, double safety, Entry *out)
{
    typedef long double L;
    const int N = DEG + 1;
    const int e = E_LO + (k >> M_BITS), j = k & ((1 << M_BITS) - 1);
    const L xk = ldexpl((L)1 + ((L)j + (L)0.5) / (L)(1 << M_BITS), e);
    const L alpha = (L)(1.0 / order) / 2;       // corr = x^(-alpha); 1/order is the DOUBLE the reference raises to
    const L ck = powl(xk, -alpha);
    // g(r) = (1 + r)^(-alpha), r = (x - xk)/xk
    L g[N + 1], u[N + 1], w[N + 1], up[N + 1], q[N + 1], f[N + 1];
    g[0] = 1;
    for (int n = 0; n < N; ++n) g[n + 1] = g[n] * (-alpha - (L)n) / (L)(n + 1);
    // u(r) = safety * ck * g(r) - 1
    for (int n = 0; n <= N; ++n) u[n] = (L)safety * ck * g[n];
    u[0] -= 1;
    // w = 1 + u^2;  up = du/dr;  q = up / w;  atan(u(r)) = atan(u0) + integral of q
    for (int n = 0; n < N; ++n) {
        L s = 0;
        for (int i = 0; i <= n; ++i) s += u[i] * u[n - i];
        w[n] = s + (n == 0 ? 1 : 0);
        up[n] = (L)(n + 1) * u[n + 1];
    }
    for (int n = 0; n < N; ++n) {
        L s = up[n];
        for (int i = 1; i <= n; ++i) s -= w[i] * q[n - i];
        q[n] = s / w[0];
    }
    f[0] = 1 + atanl(u[0]);
    for (int n = 1; n < N; ++n) f[n] = q[n - 1] / (L)n;
    L scale = 1;
    for (int n = 0; n < N; ++n) {
        out->c[n] = (double)(f[n] * scale);      // coefficient of d^n = coefficient of r^n / xk^n
        scale /= xk;
    }
}

inline void build_table(double order, double safety, Entry *tab /* [ENTRIES] */)
{
    for (int k = 0; k < ENTRIES; ++k) build_entry(k, order, safety, &tab[k]);
}

// ---- evaluation (host + device) -----------------------------------------------------------------------------
#if defined(__CUDACC__)
#define DSB_FT_HD __host__ __device__ __forceinline__
#else
#define DSB_FT_HD inline
#endif

// table index of x (any bit pattern); valid iff (unsigned) result < ENTRIES
DSB_FT_HD int index_of_hi(uint32_t hi) { return (int)(hi >> INDEX_SHIFT) - INDEX_BASE; }

// interval midpoint from the high word of x: keep sign/exponent/top M_BITS mantissa bits, set the next bit
DSB_FT_HD uint32_t midpoint_hi(uint32_t hi) { return (hi & ~((1u << INDEX_SHIFT) - 1u)) | (1u << (INDEX_SHIFT - 1)); }

DSB_FT_HD double horner(const double (&c)[DEG + 1], double d)
{
    double p = c[DEG];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int n = DEG - 1; n >= 0; --n) p = fma(p, d, c[n]);
    return p;
}

#if !defined(__CUDACC__)
// host reference evaluation (tests): F(x) through the table, or NaN when x is outside it
inline double eval_host(const Entry *tab, double x)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    uint32_t hi = (uint32_t)(b >> 32);
    int k = index_of_hi(hi);
    if ((unsigned)k >= (unsigned)ENTRIES) return NAN;
    uint64_t mb = (uint64_t)midpoint_hi(hi) << 32;
    double xk;
    memcpy(&xk, &mb, 8);
    return horner(tab[k].c, x - xk);
}
#endif

}  // namespace ft
}  // namespace dsb
