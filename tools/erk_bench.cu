// erk_bench.cu -- standalone harness for tuning the fused ERK kernel: builds ONE instantiation with the
// launch bounds given by -DDSB_ERK_THREADS / -DDSB_ERK_MIN_BLOCKS and times it on a 2^20 Lorenz ensemble.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -DDSB_ERK_MIN_BLOCKS=5 -o erk_bench tools/erk_bench.cu
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include <random>

#include "../desolver_b200/csrc/plan.h"
#include "../desolver_b200/csrc/rhs_builtin.cuh"

namespace dsb { void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); } }
using namespace dsb;

int main(int argc, char **argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : (1 << 20);
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    using Tab = StaticTab<CashKarp45>;
    using Rhs = LorenzRhs<false>;
    std::mt19937_64 rng(0);
    std::uniform_real_distribution<double> U(0, 1);
    std::vector<double> y0(3 * n), t0(n, 0.0), dt0(n, 1e-2);
    for (long long i = 0; i < n; ++i) { y0[i] = -20 + 40 * U(rng); y0[n + i] = -25 + 50 * U(rng); y0[2 * n + i] = 50 * U(rng); }
    double *y, *t, *dt, *prm; int *acc, *rej, *st; unsigned long long *ctr; double2 *tab;
    cudaMalloc(&y, 24 * n); cudaMalloc(&t, 8 * n); cudaMalloc(&dt, 8 * n); cudaMalloc(&prm, 24);
    cudaMalloc(&acc, 4 * n); cudaMalloc(&rej, 4 * n); cudaMalloc(&st, 4 * n); cudaMalloc(&ctr, 64); cudaMalloc(&tab, sizeof(double2) * fm::ATAN_TABLE_SIZE);
    double hp[3] = {10.0, 28.0, 8.0 / 3.0};
    cudaMemcpy(prm, hp, 24, cudaMemcpyHostToDevice);
    double2 htab[fm::ATAN_TABLE_SIZE]; fm::build_atan_table(htab);
    cudaMemcpy(tab, htab, sizeof(htab), cudaMemcpyHostToDevice);

    dsb_erk_plan plan;
    plan.stages = 6; plan.final_rows = 2; plan.order = 5.0;
    plan.inter.assign(42, 0.0); plan.fin.assign(14, 0.0);
    for (int s = 0; s < 6; ++s) { plan.inter[s * 7] = CashKarp45::C[s]; for (int j = 0; j < 6; ++j) plan.inter[s * 7 + 1 + j] = CashKarp45::A[s][j];
                                  plan.fin[1 + s] = CashKarp45::B[s]; plan.fin[8 + s] = CashKarp45::BH[s]; }
    ErkKernelArgs<6> args{};
    fill_tableau<6>(&plan, args.tab);
    args.atan_table = tab;
    std::vector<ft::Entry> hft(ft::ENTRIES); ft::build_table(5.0, 0.8, hft.data());
    ft::Entry *dft; cudaMalloc(&dft, sizeof(ft::Entry) * ft::ENTRIES); cudaMemcpy(dft, hft.data(), sizeof(ft::Entry) * ft::ENTRIES, cudaMemcpyHostToDevice);
    args.factor_table = dft;
    dsb_erk_run_args &a = args.a;
    a.n = n; a.y = y; a.t = t; a.dt = dt;
    for (int p = 0; p < 3; ++p) { a.params[p] = prm + p; a.param_stride[p] = 0; }
    a.tf = 2.0; a.rtol = 1e-9; a.atol = 1e-9; a.accepted = acc; a.rejected = rej; a.status = st;
    a.max_steps = -1; a.first_call = 1; a.counters = (uint64_t *)ctr;
#ifdef BENCH_FRESH
    a.fresh = 1; a.t_init = 0.0; a.dt_init = 1e-2;
#endif

    int per_sm = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    occupancy_erk_fused<Rhs, Tab, false>(&per_sm);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, erk_fused_kernel<Rhs, Tab, false>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; long long attempts = 0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaMemcpy(y, y0.data(), 24 * n, cudaMemcpyHostToDevice); cudaMemcpy(t, t0.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(dt, dt0.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaMemset(acc, 0, 4 * n); cudaMemset(rej, 0, 4 * n); cudaMemset(ctr, 0, 64);
        cudaEventRecord(e0);
        #ifdef BENCH_FLOW
        launch_erk_fused<Rhs, Tab, false, true>(args, sms * per_sm, 0);     // flow-controlled instantiation, hooks idle
#else
        launch_erk_fused<Rhs, Tab, false>(args, sms * per_sm, 0);
#endif
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    std::vector<int> ha(n), hr(n);
    cudaMemcpy(ha.data(), acc, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(hr.data(), rej, 4 * n, cudaMemcpyDeviceToHost);
    for (long long i = 0; i < n; ++i) attempts += ha[i] + hr[i];
    {   // checksums of the result (two builds of the kernel must agree bit for bit) and the kernel's own statistics
        std::vector<double> hy(3 * n), ht(n);
        std::vector<int> hs(n);
        unsigned long long hc[8];
        cudaMemcpy(hy.data(), y, 24 * n, cudaMemcpyDeviceToHost); cudaMemcpy(ht.data(), t, 8 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hs.data(), st, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(hc, ctr, 64, cudaMemcpyDeviceToHost);
        unsigned long long x = 0, sa = 0, sr = 0, ss = 0;
        for (long long i = 0; i < 3 * n; ++i) { unsigned long long b; memcpy(&b, &hy[i], 8); x = x * 1099511628211ull + b; }
        for (long long i = 0; i < n; ++i) { unsigned long long b; memcpy(&b, &ht[i], 8); x = x * 1099511628211ull + b; sa += ha[i]; sr += hr[i]; ss += hs[i]; }
        printf("checksum %016llx accepted %llu rejected %llu status-sum %llu | counters: attempts %llu accepted %llu failed %llu running %llu\n",
               x, sa, sr, ss, hc[2], hc[3], hc[4], hc[5]);
    }
    printf("threads %d min_blocks %d regs %d blocks/SM %d : best %.3f ms  %.3e traj-steps/s  (%.1f%% of 8.17e10)  err=%s\n",
           ERK_THREADS, erk_min_blocks<Rhs, 6>(), fa.numRegs, per_sm, best, attempts / (best * 1e-3), 100.0 * attempts / (best * 1e-3) / 8.17e10,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
