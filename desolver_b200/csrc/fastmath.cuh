// fastmath.cuh -- double-precision building blocks of the fused step-size controller.
//
// The reference controller (integrators/integrator_template.py:62-93) needs, per attempt and
// per trajectory: dim divisions, a sqrt, a reciprocal, pow(x, 1/order) and atan.  With libm
// those are ~60 % of the FP64 instructions of a Lorenz RK45 attempt, and the FP64 pipe is the
// co-limit of the kernel.  These replacements keep <= ~2 ulp accuracy (so the
// committed dt differs from libm's by the same last-ulp amount that numpy-SIMD and glibc differ
// from each other) at a fraction of the instruction count:
//   fast_div      : MUFU.RCP64H seed + one Newton step + remainder correction     (5 DFMA)
//   inv_root<P>   : x^(-1/P) = (||e||^2)^(-1/(2 order)) -- folds sqrt, reciprocal and pow into
//                   an FP32 lg2/ex2 seed + two division-free Newton steps          (14 DFMA)
//   one_plus_atan : table-driven argument reduction + 4-term series                (~15 DFMA)
// Everything is __host__ __device__ so tests/test_fastmath.py can check the algorithms on the
// CPU against libm (the MUFU seeds are replaced by deliberately truncated libm values there).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DSB_HD __host__ __device__ __forceinline__
#else
#define DSB_HD inline
struct double2 { double x, y; };
#endif

namespace dsb {
namespace fm {

// Double literals of the hot loop live in the user constant bank: a constant-bank operand is free, while a
// 64-bit immediate that does not fit the 32-bit immediate field costs two UMOVs every time it is used.
// (A __constant__ object may be rewritten by the host, so the compiler cannot fold it back into immediates.)
struct HotConstants {
    double c08, c02, ip10, hp10, ip16, hp16, am7, a5, am3, reject_below, eps32, xlo, xhi, ubig;
};
#if defined(__CUDACC__)
static __constant__ HotConstants kc = {0.8, 0.2, 1.0 / 10, 11.0 / 20, 1.0 / 16, 17.0 / 32, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0,
                                       0.9 * 0.9, 32.0 * 2.220446049250313e-16, 1e-36, 1e36, 1e30};
#endif
#if defined(__CUDA_ARCH__)
#define DSB_KC(name, literal) (::dsb::fm::kc.name)
#else
#define DSB_KC(name, literal) (literal)
#endif

constexpr int ATAN_TABLE_SIZE = 128;
constexpr double ATAN_G_LO = -0.5;
constexpr double ATAN_INV_DG = 128.0 / 1.5;   // table is uniform in g = u / (1 + |u|), g in [-0.5, 1)

DSB_HD uint64_t as_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return b;
#endif
}

DSB_HD float bits_as_float(uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// ~2^-20-accurate reciprocal seed (device: one MUFU.RCP64H; host: libm truncated to match)
DSB_HD double rcp_seed(double d)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
#else
    uint64_t b = as_bits(1.0 / d) & 0xffffffff00000000ull;
    double r; memcpy(&r, &b, 8); return r;
#endif
}

DSB_HD float lg2_seed(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return log2f(x) * (1.0f + 3e-7f);
#endif
}

DSB_HD float ex2_seed(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x) * (1.0f - 3e-7f);
#endif
}

DSB_HD float rcpf_seed(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// max(|a|, |b|) through the integer pipe: for non-negative IEEE doubles the bit patterns order like
// the values (a NaN, having the largest exponent, wins -- np.maximum semantics).  Saves the FP64
// pipe a DSETP plus the select/NaN-quieting sequence fmax() expands to on sm_100.
DSB_HD double abs_max(double a, double b)
{
#if defined(__CUDA_ARCH__)
    // explicit 32-bit halves: written as a 64-bit mask the compiler materialises |x| with a DADD (FP64 pipe)
    unsigned ha = (unsigned)__double2hiint(a) & 0x7fffffffu, hb = (unsigned)__double2hiint(b) & 0x7fffffffu;
    unsigned la = (unsigned)__double2loint(a), lb = (unsigned)__double2loint(b);
    bool a_wins = (ha > hb) || (ha == hb && la > lb);
    return __hiloint2double((int)(a_wins ? ha : hb), (int)(a_wins ? la : lb));
#else
    uint64_t ia = as_bits(a) & 0x7fffffffffffffffull, ib = as_bits(b) & 0x7fffffffffffffffull;
    uint64_t m = ia > ib ? ia : ib;
    double r; memcpy(&r, &m, 8); return r;
#endif
}

// max(|a|, |b|) with ONE FP64 instruction: DSETP takes |.| operand modifiers for free, the winner is
// picked with two 32-bit selects and its sign bit cleared with one LOP -- 2 + 3 issue slots against the 8 of
// the all-integer version above (an FP64 instruction occupies its sub-partition's issue port for two cycles).
// NaN handling differs from np.maximum only when exactly one operand is NaN (the comparison is false, b wins);
// in the kernel a NaN state makes the error estimate NaN as well, so the attempt is rejected either way.
DSB_HD double abs_max_setp(double a, double b)
{
#if defined(__CUDA_ARCH__)
    const bool a_wins = fabs(a) > fabs(b);
    const int hi = a_wins ? __double2hiint(a) : __double2hiint(b);
    const int lo = a_wins ? __double2loint(a) : __double2loint(b);
    return __hiloint2double(hi & 0x7fffffff, lo);
#else
    return fabs(a) > fabs(b) ? fabs(a) : fabs(b);
#endif
}

// n / d to ~1 ulp for normal d; d = 0, inf, nan give inf/nan like IEEE division does.
DSB_HD double fast_div(double n, double d)
{
    double r = rcp_seed(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);               // 2^-40
    double q = n * r;
    double rem = fma(-d, q, n);
    return fma(rem, r, q);
}

// n / d to ~2^-40 relative (seed + one Newton step, no remainder correction).  Used ONLY for the error-norm
// quotients err/tol: the norm feeds nothing but the step-size factor, where 1e-12 relative is far below the
// conditioning of dt itself (a last-ulp state change already moves dt by ~1e-9 relative).
DSB_HD double quick_div(double n, double d)
{
    double r = rcp_seed(d);
    r = fma(r, fma(-d, r, 1.0), r);
    return n * r;
}

template <int P>
DSB_HD double ipow(double c)
{
    if (P == 1) return c;
    double h = ipow<(P / 2 > 0 ? P / 2 : 1)>(c);
    double sq = h * h;
    return (P & 1) ? sq * c : sq;
}

DSB_HD double ipow_rt(double c, int p)
{
    double r = 1.0;
    while (p > 0) {
        if (p & 1) r *= c;
        c *= c;
        p >>= 1;
    }
    return r;
}

// FP32 seed of x^(-1/P) for normal positive x, via exponent/mantissa split so any double
// magnitude is safe (the seed itself always fits a float: |log2 x| / P <= 1074 / 2).
DSB_HD double inv_root_seed(double x, float inv_p)
{
    uint64_t b = as_bits(x);
    uint32_t hi = (uint32_t)(b >> 32);
    int e = (int)((hi >> 20) & 0x7ffu) - 1023;
    uint32_t mant23 = ((hi & 0xfffffu) << 3) | ((uint32_t)(b >> 29) & 7u);
    float m = bits_as_float(0x3f800000u | mant23);         // [1, 2)
    float lg = (float)e + lg2_seed(m);
    return (double)ex2_seed(-lg * inv_p);
}

// x^(-1/P) for x in [1e-280, 1e280]: ONE third-order (Halley-type) correction of the FP32 seed,
//   e = 1 - x c^P,   c <- c + c * (e/P) * (1 + (P+1)/(2P) e),
// whose error is O(e^3): a 3e-7 seed lands below 1e-17 relative before rounding.
template <int P>
DSB_HD double inv_root(double x)
{
    double c = inv_root_seed(x, 1.0f / (float)P);
    double e = fma(-x, ipow<P>(c), 1.0);
    const double ip = P == 10 ? DSB_KC(ip10, 0.1) : (P == 16 ? DSB_KC(ip16, 1.0 / 16) : 1.0 / P);
    const double hp = P == 10 ? DSB_KC(hp10, 11.0 / 20) : (P == 16 ? DSB_KC(hp16, 17.0 / 32) : (P + 1.0) / (2.0 * P));
    double g = fma(e, hp, 1.0) * (e * ip);
    return fma(c, g, c);
}

// Same, for x in [1e-36, 1e36] (|err/tol| in [1e-18, 1e18]): the seed is taken from (float)x directly
// (one F2F instead of the exponent/mantissa split).
template <int P>
DSB_HD double inv_root_narrow(double x)
{
    double c = (double)ex2_seed(lg2_seed((float)x) * (-1.0f / (float)P));
    double e = fma(-x, ipow<P>(c), 1.0);
    const double ip = P == 10 ? DSB_KC(ip10, 0.1) : (P == 16 ? DSB_KC(ip16, 1.0 / 16) : 1.0 / P);
    const double hp = P == 10 ? DSB_KC(hp10, 11.0 / 20) : (P == 16 ? DSB_KC(hp16, 17.0 / 32) : (P + 1.0) / (2.0 * P));
    double g = fma(e, hp, 1.0) * (e * ip);
    return fma(c, g, c);
}

DSB_HD double inv_root_rt(double x, int p)
{
    double c = inv_root_seed(x, 1.0f / (float)p);
    double ip = 1.0 / (double)p;
    double e = fma(-x, ipow_rt(c, p), 1.0);
    double g = fma(e, 0.5 * (p + 1.0) * ip, 1.0) * (e * ip);
    return fma(c, g, c);
}

// The reference's 3-term controller product for the 3rd and later attempt of a step
// (integrators/integrator_template.py:80-90), for finite positive squared norms x = ||e/tol||^2 (eps = x^-1/2):
//   eps_cur^(0.55/order) * eps_last^(-0.27/order) * eps_last_last^(0.05/order)
//     = x_cur^(-11/(40 order)) * x_last^(27/(200 order)) * x_ll^(-1/(40 order))
// written with integer inverse roots (q = 40*order, 5q = 200*order must be integers) so that no libm pow is
// needed.  The 200*order-th root is refined twice (its third-order constant is ~P^2/3).
DSB_HD double corr_three_term(double x_cur, double x_last, double x_ll, int q)
{
    double r1 = inv_root_rt(x_cur, q);
    double r3 = inv_root_rt(x_ll, q);
    const int p2 = 5 * q;
    double c = inv_root_seed(x_last, 1.0f / (float)p2);
    const double ip = 1.0 / (double)p2;
    for (int it = 0; it < 2; ++it) {
        double e = fma(-x_last, ipow_rt(c, p2), 1.0);
        c = fma(c, fma(e, 0.5 * (p2 + 1.0) * ip, 1.0) * (e * ip), c);
    }
    return (ipow_rt(r1, 11) * fast_div(1.0, ipow_rt(c, 27))) * r3;
}

// Host-side table builder: entry k = { u_k, 1 + atan(u_k) } with g_k = -0.5 + 1.5 k / 128.
inline void build_atan_table(double2 *tab)
{
    for (int k = 0; k < ATAN_TABLE_SIZE; ++k) {
        double g = ATAN_G_LO + (double)k / ATAN_INV_DG;
        double u = g / (1.0 - fabs(g));
        tab[k].x = u;
        tab[k].y = 1.0 + atan(u);
    }
}

// 1 + atan(u) for u in [-1, 1e30): atan(u) = atan(u_k) + atan((u - u_k) / (1 + u u_k)).
DSB_HD double one_plus_atan(double u, const double2 *tab)
{
    float uf = fminf((float)u, 1e18f);           // u up to 1e300 (corr cap): (float)u = inf would make g NaN
    float g = uf * rcpf_seed(1.0f + fabsf(uf));
    int k = (int)rintf((g - (float)ATAN_G_LO) * (float)ATAN_INV_DG);
    k = k < 0 ? 0 : (k > ATAN_TABLE_SIZE - 1 ? ATAN_TABLE_SIZE - 1 : k);
    double2 e = tab[k];
    double r = fast_div(u - e.x, fma(u, e.x, 1.0));
    double r2 = r * r;
    double p = fma(r2, fma(r2, DSB_KC(am7, -1.0 / 7.0), DSB_KC(a5, 1.0 / 5.0)), DSB_KC(am3, -1.0 / 3.0));   // |r| < 0.012: r^9/9 < 6e-19
    return e.y + fma(r * r2, p, r);
}

}  // namespace fm
}  // namespace dsb
