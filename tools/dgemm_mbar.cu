// EXPERIMENT (tools/dgemm_sweep.py): the FP64 DMMA GEMM of csrc/dgemm.cu with the per-k-step CTA barrier replaced by
// mbarriers -- full[s] (all 128 threads' cp.async of stage s have landed: cp.async.mbarrier.arrive.noinc) and empty[s]
// (all 4 warps have finished reading stage s) -- over a 4-stage ring with a prefetch distance of 2, so that the warps of a
// CTA may drift apart by one k-step instead of meeting at a barrier every 16 columns of K.  Same tiles, same shared-memory
// layout, same arithmetic order as dgemm_dmma_kernel (results must be bit-identical).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mb {
constexpr int BM = 64, BN = 128, BK = 16, S = 4, PD = 2, THREADS = 128, CTAS = 2, NT = 8;
constexpr int LDA = BK + 4, LDB = BN + 8;
constexpr int A_STAGE = BM * LDA, B_STAGE = BK * LDB;
constexpr int SMEM_BYTES = S * (A_STAGE + B_STAGE) * (int)sizeof(double);

__device__ __forceinline__ unsigned sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sptr(smem)), "l"(gmem));
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t *b)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(sptr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
}  // namespace mb

__global__ void __launch_bounds__(mb::THREADS, mb::CTAS)
dgemm_mbar_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int M, int N, int K)
{
    using namespace mb;
    extern __shared__ __align__(16) double smem[];
    __shared__ uint64_t full[S], empty[S];
    double *sA = smem, *sB = smem + S * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % 2, wn = warp / 2;
    const int g = lane >> 2, t = lane & 3;
    const int mbk = M / BM, nbk = N / BN, kt = K / BK;
    const int tiles = mbk * nbk;
    if (tid == 0)
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], THREADS); mbar_init(&empty[s], THREADS / 32); }
    __syncthreads();
    const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const long long total = (long long)my_tiles * kt;

    // load stream state: global step gl = (tile index li, k-step lk)
    long long gl = 0;
    int li = 0, lk = 0;
    auto issue = [&]() {
        const int tile = blockIdx.x + li * gridDim.x;
        const int tm = tile % mbk, tn = tile / mbk;
        const double *Ag = A + (size_t)(tm * BM) * K + lk * BK;
        const double *Bg = B + (size_t)(lk * BK) * N + tn * BN;
        const int s = (int)(gl % S);
        const long long use = gl / S;
        if (use > 0) mbar_wait(&empty[s], (unsigned)((use - 1) & 1));
        double *a = sA + s * A_STAGE, *b = sB + s * B_STAGE;
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / THREADS; ++i) {
            const int c = tid + i * THREADS, row = c / (BK / 2), col = (c % (BK / 2)) * 2;
            cp_async16(a + row * LDA + col, Ag + (size_t)row * K + col);
        }
#pragma unroll
        for (int i = 0; i < (BK * BN / 2) / THREADS; ++i) {
            const int c = tid + i * THREADS, row = c >> 6, col = (c & 63) * 2;
            cp_async16(b + row * LDB + col, Bg + (size_t)row * N + col);
        }
        mbar_arrive_cp_async(&full[s]);
        ++gl;
        if (++lk == kt) { lk = 0; ++li; }
    };
    for (int p = 0; p < PD; ++p)
        if (gl < total) issue();

    long long gc = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int tile = blockIdx.x + ti * gridDim.x;
        const int m0 = (tile % mbk) * BM, n0 = (tile / mbk) * BN;
        double acc[4][NT][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int k = 0; k < kt; ++k, ++gc) {
            if (gl < total) issue();                                     // step gc + PD
            const int s = (int)(gc % S);
            mbar_wait(&full[s], (unsigned)((gc / S) & 1));
            const double *a = sA + s * A_STAGE + (wm * 32 + g) * LDA + t;
            const double *b = sB + s * B_STAGE + t * LDB + wn * 64 + g;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double af[4], bf[NT];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = a[i * 8 * LDA + kk * 4];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = b[kk * 4 * LDB + j * 8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc[i][j], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        double *Cg = C + (size_t)(m0 + wm * 32 + g) * N + n0 + wn * 64 + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
                *reinterpret_cast<double2 *>(Cg + (size_t)i * 8 * N + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

extern "C" __attribute__((visibility("default"))) int variant_dgemm(const double *A, const double *B, double *C, long long M,
                                                                    long long N, long long K, int sms, void *stream)
{
    using namespace mb;
    if (M % BM || N % BN || K % BK) return 3;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dgemm_mbar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return 2;
        configured = true;
    }
    const long long tiles = (M / BM) * (N / BN);
    const int grid = (int)(tiles < (long long)sms * CTAS ? tiles : (long long)sms * CTAS);
    dgemm_mbar_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A, B, C, (int)M, (int)N, (int)K);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
