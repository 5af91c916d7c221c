// Host-side accuracy check of desolver_b200/csrc/fastmath.cuh (compiled by tests/test_fastmath.py with g++).
// The MUFU seeds are replaced by truncated/perturbed libm values of the same accuracy class.
#include "../desolver_b200/csrc/fastmath.cuh"
#include <stdio.h>
#include <random>
using namespace dsb::fm;
int main()
{
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0, 1);
    double2 tab[ATAN_TABLE_SIZE];
    build_atan_table(tab);
    double w10 = 0, w16 = 0, wrt = 0, wat = 0, wdiv = 0, wmax = 0;
    for (int i = 0; i < 2000000; i++) {
        double x = (i % 2) ? pow(10.0, -3 + 6 * U(rng)) : pow(10.0, -8 + 16 * U(rng));
        double r10 = pow(x, -0.1), r16 = pow(x, -1.0 / 16);
        w10 = fmax(w10, fabs(inv_root<10>(x) - r10) / r10);
        w16 = fmax(w16, fabs(inv_root<16>(x) - r16) / r16);
        w10 = fmax(w10, fabs(inv_root_narrow<10>(x) - r10) / r10);
        wrt = fmax(wrt, fabs(inv_root_rt(x, 10) - r10) / r10);
        double u = (i % 3 == 0) ? -1 + 4 * U(rng) : ((i % 3 == 1) ? -1 + pow(10.0, -8 + 36 * U(rng)) : 0.8 * pow(10.0, -1 + 2 * U(rng)) - 1);
        if (u < -1) u = -1;
        double fr = 1 + atan(u);
        wat = fmax(wat, fabs(one_plus_atan(u, tab) - fr) / fr);
        double n = U(rng) - 0.5, d = pow(10.0, -12 + 24 * U(rng));
        wdiv = fmax(wdiv, fabs(fast_div(n, d) - n / d) / fabs(n / d));
        double a = U(rng) - 0.5, b = (U(rng) - 0.5) * pow(10.0, -3 + 6 * U(rng));
        wmax = fmax(wmax, fabs(abs_max(a, b) - fmax(fabs(a), fabs(b))));
    }
    // extreme magnitudes stay accurate up to the inexactness of the DOUBLE 0.2 the reference raises to
    double wext = 0;
    for (int i = 0; i < 200000; i++) {
        double x = pow(10.0, -280 + 560 * U(rng));
        wext = fmax(wext, fabs(inv_root<16>(x) - pow(x, -1.0 / 16)) / pow(x, -1.0 / 16));
    }
    // 3-term controller product against libm pow, orders 5 and 8
    double w3 = 0;
    for (int i = 0; i < 300000; i++) {
        double a = pow(10.0, -6 + 12 * U(rng)), b = pow(10.0, -6 + 12 * U(rng)), c = pow(10.0, -6 + 12 * U(rng));
        for (int order = 5; order <= 8; order += 3) {
            double ref = pow(1 / sqrt(a), 0.55 / order) * pow(1 / sqrt(b), -0.27 / order) * pow(1 / sqrt(c), 0.05 / order);
            w3 = fmax(w3, fabs(corr_three_term(a, b, c, 40 * order) - ref) / ref);
        }
    }
    printf("%.3e %.3e %.3e %.3e %.3e %.3e %.3e %.3e\n", w10, w16, wrt, wat, wdiv, wmax, wext, w3);
    printf("%.17g %.17g\n", one_plus_atan(-1.0, tab), one_plus_atan(0.8e300, tab));
    return 0;
}
