// dgemm.cu -- K3: the dense linear right-hand side  K = A @ Y_stage  over the ensemble as a hand-written FP64
// tensor-core GEMM (sm_100a).
//
// Replaces the user's `A @ y` of the linear benchmark system (rhs_plugins.linear; the reference evaluates it with
// numpy/OpenBLAS inside compute_step, integrators/components/runge_kutta_methods.py:49-54) for an ensemble stored as
// Y[n][columns] (trajectory index fastest): C[M][N] = A[M][K] * B[K][N], all row-major fp64.
//
// Blackwell's tcgen05 MMA has no f64 kind, so FP64 tensor work is the warp-level DMMA `mma.sync.m8n8k4.f64`
// (256 FMA per warp-instruction; the FP64 pipe retires 64 FMA/clk/SM, i.e. one DMMA per 16 cycles per SM
// sub-partition).  The kernel is therefore bound by that pipe as long as operands arrive:
//   * persistent CTAs, TWO per SM with 4 warps each (64 x 128 output tiles, static round-robin, m fastest so that
//     the CTAs of one wave share B column blocks and stream A once): while one CTA sits in its per-k-step barrier
//     or writes its results, the other keeps the pipe busy (one 8-warp CTA with 128 x 128 tiles: 8.47 ms;
//     two 4-warp CTAs: 8.02 ms; cuBLAS 7.78 ms; the pipe's own limit 7.4 ms);
//   * K in steps of 16 through a 3-stage cp.async pipeline (16-byte copies, 27 KB per stage); the next tile's
//     first stages go in flight before the current tile's results are stored;
//   * shared-memory leading dimensions 20 (A, k fastest) and 136 (B, n fastest) doubles: the fragment loads of a
//     warp (A[g][t], B[t][g], g = lane/4, t = lane%4) hit every bank group exactly once per wavefront;
//   * warp tile 32 x 64: 4 + 8 fragment loads feed 32 DMMAs per k-step of 4; 64 accumulator doubles per lane
//     (246 registers, which is what limits the SM to two CTAs).
// Shapes with M % 64 == 0, N % 128 == 0, K % 16 == 0 (the benchmark configuration 4096 x 4096 x 8192 and every
// power-of-two problem) run the unpredicated instantiation; any other shape with even N and K runs the EDGE instantiation
// (zero-filled cp.async for chunks outside the matrices, guarded stores).  Odd N or K: DSB_ERR_UNSUPPORTED (library GEMM).
// Results are bit-identical to cuBLAS DGEMM on the shapes tested (same k order per accumulator).
#include "common.cuh"

namespace dsb {

namespace gemm {
#ifndef DSB_GEMM_BK
#define DSB_GEMM_BK 16
#endif
#ifndef DSB_GEMM_STAGES
#define DSB_GEMM_STAGES 3
#endif
#ifndef DSB_GEMM_WM
#define DSB_GEMM_WM 2          // warps along M: CTA tile (32 * WM) x 128, 2 * WM warps
#endif
#ifndef DSB_GEMM_CTAS
#define DSB_GEMM_CTAS 2        // resident CTAs per SM
#endif
#ifndef DSB_GEMM_WTN
#define DSB_GEMM_WTN 64        // warp tile width (columns): 64 -> 2 warps along N, 32 -> 4 warps along N
#endif
#ifndef DSB_GEMM_GROUP_M
#define DSB_GEMM_GROUP_M 0     // > 0: tiles are walked in groups of GROUP_M row blocks (squarer set of concurrent tiles)
#endif
#ifndef DSB_GEMM_FRAG_DB
#define DSB_GEMM_FRAG_DB 0     // 1: fragments of k-substep kk+1 are loaded from shared memory before the DMMAs of kk issue
#endif
constexpr int WM = DSB_GEMM_WM, CTAS = DSB_GEMM_CTAS, WTN = DSB_GEMM_WTN, NT = WTN / 8, GROUP_M = DSB_GEMM_GROUP_M;
constexpr int BM = 32 * WM, BN = 128, BK = DSB_GEMM_BK, STAGES = DSB_GEMM_STAGES, THREADS = 32 * WM * (BN / WTN);
constexpr int LDA = BK + 4;        // doubles: 2 * LDA words == 8 (mod 32) -> rows g = 0..3 fill one wavefront
#ifndef DSB_GEMM_LDB_PAD
#define DSB_GEMM_LDB_PAD 8
#endif
constexpr int LDB = BN + DSB_GEMM_LDB_PAD;   // doubles; the k rows t = 0..3 of a half-warp must land in different bank groups
constexpr int A_STAGE = BM * LDA, B_STAGE = BK * LDB;
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
// the same with `bytes` (0 or 16) taken from global memory and the rest of the 16 bytes zero-filled: edge tiles
__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, unsigned bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
}  // namespace gemm

// EDGE: M, N, K need not be multiples of the tile (N and K even, so that every 16-byte chunk is inside or outside as a
// whole): chunks outside the matrices are zero-filled by cp.async, stores outside C are skipped.  A separate
// instantiation, so that the aligned shapes of the benchmark keep their unpredicated loads.
template <bool EDGE>
__global__ void __launch_bounds__(gemm::THREADS, gemm::CTAS)
dgemm_dmma_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int M, int N, int K)
{
    using namespace gemm;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem, *sB = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;                // WM x 2 warps -> warp tile 32 x 64
    const int g = lane >> 2, t = lane & 3;
    const int mb = EDGE ? (M + BM - 1) / BM : M / BM, nb = EDGE ? (N + BN - 1) / BN : N / BN;
    const int kt = EDGE ? (K + BK - 1) / BK : K / BK;

    // tile index -> (row block, column block)
    auto tile_mn = [&](int tile, int &tm, int &tn) {
        if (GROUP_M > 0) {
            const int per_group = GROUP_M * nb, grp = tile / per_group, first = grp * GROUP_M;
            const int rows = (mb - first) < GROUP_M ? (mb - first) : GROUP_M, r = tile - grp * per_group;
            tm = first + r % rows;
            tn = r / rows;
        } else {
            tm = tile % mb;
            tn = tile / mb;
        }
    };
    const double *Ag = nullptr, *Bg = nullptr;
    int m_load = 0, n_load = 0;                                          // first row / column of the tile being loaded
    auto load_stage = [&](int stage, int k0) {
        double *a = sA + stage * A_STAGE, *b = sB + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / THREADS; ++i) {              // A tile: 128 rows x BK/2 chunks of 2 doubles
            const int c = tid + i * THREADS, row = c / (BK / 2), col = (c % (BK / 2)) * 2;
            if (EDGE) {
                const bool in = m_load + row < M && k0 + col < K;
                cp_async16_zfill(a + row * LDA + col, in ? Ag + (size_t)row * K + k0 + col : A, in ? 16u : 0u);
            } else {
                cp_async16(a + row * LDA + col, Ag + (size_t)row * K + k0 + col);
            }
        }
#pragma unroll
        for (int i = 0; i < (BK * BN / 2) / THREADS; ++i) {              // B tile: BK rows x 64 chunks
            const int c = tid + i * THREADS, row = c >> 6, col = (c & 63) * 2;
            if (EDGE) {
                const bool in = k0 + row < K && n_load + col < N;
                cp_async16_zfill(b + row * LDB + col, in ? Bg + (size_t)(k0 + row) * N + col : B, in ? 16u : 0u);
            } else {
                cp_async16(b + row * LDB + col, Bg + (size_t)(k0 + row) * N + col);
            }
        }
    };
    // prologue of a tile: the first STAGES-1 k-steps in flight
    auto prologue = [&](int tile) {
        int tm, tn;
        tile_mn(tile, tm, tn);
        m_load = tm * BM;
        n_load = tn * BN;
        Ag = A + (size_t)(tm * BM) * K;
        Bg = B + tn * BN;
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < kt) load_stage(s, s * BK);
            cp_async_commit();
        }
    };

    int tile = blockIdx.x;
    if (tile < mb * nb) prologue(tile);
    for (; tile < mb * nb; tile += gridDim.x) {
        int tm0, tn0;
        tile_mn(tile, tm0, tn0);
        const int m0 = tm0 * BM, n0 = tn0 * BN;
        double acc[4][NT][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int k = 0; k < kt; ++k) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();                                             // stage k landed; stage k-1 is free again
            if (k + STAGES - 1 < kt) load_stage((k + STAGES - 1) % STAGES, (k + STAGES - 1) * BK);
            cp_async_commit();
            const double *a = sA + (k % STAGES) * A_STAGE + (wm * 32 + g) * LDA + t;
            const double *b = sB + (k % STAGES) * B_STAGE + t * LDB + wn * WTN + g;
#if DSB_GEMM_FRAG_DB
            double af[2][4], bf[2][NT];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[0][i] = a[i * 8 * LDA];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[0][j] = b[j * 8];
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                if (kk + 1 < BK / 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[(kk + 1) & 1][i] = a[i * 8 * LDA + (kk + 1) * 4];
#pragma unroll
                    for (int j = 0; j < NT; ++j) bf[(kk + 1) & 1][j] = b[(kk + 1) * 4 * LDB + j * 8];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc[i][j], af[kk & 1][i], bf[kk & 1][j]);
            }
#else
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
#endif
        }
        cp_async_wait<0>();
        __syncthreads();                                                 // every warp is done with the last stages
        // the next tile's first k-steps go in flight before this tile's results are written out
        // (requires kt >= STAGES - 1 stages per tile to keep the stage indices of the pipeline aligned: kt % STAGES
        // may differ, so the next tile simply restarts at stage 0 -- all stages are free here)
        if (tile + (int)gridDim.x < mb * nb) prologue(tile + gridDim.x);

        // epilogue: lane (g, t) owns C[8 i + g][8 j + 2 t .. + 1] of each 8 x 8 tile: 16-byte stores
        const int r0 = m0 + wm * 32 + g, c0 = n0 + wn * WTN + 2 * t;
        double *Cg = C + (size_t)r0 * N + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
                if (!EDGE || (r0 + i * 8 < M && c0 + j * 8 < N))
                    *reinterpret_cast<double2 *>(Cg + (size_t)i * 8 * N + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

int launch_dgemm(const double *A, const double *B, double *C, int64_t M, int64_t N, int64_t K, int sm_count, cudaStream_t stream)
{
    using namespace gemm;
    if (M <= 0 || N <= 0 || K <= 0 || N % 2 || K % 2 || M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff ||
        ((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) % 16) {
        set_error("dsb_dgemm: needs even N and K (16-byte rows) and 16-byte aligned pointers (got %lld x %lld x %lld)",
                  (long long)M, (long long)N, (long long)K);
        return DSB_ERR_UNSUPPORTED;
    }
    const bool edge = M % BM || N % BN || K % BK;
    static bool configured[64];
    int device = 0;
    DSB_CUDA_CHECK(cudaGetDevice(&device));
    if (!configured[device & 63]) {
        DSB_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DSB_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured[device & 63] = true;
    }
    const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = (int)(tiles < (int64_t)sm_count * CTAS ? tiles : (int64_t)sm_count * CTAS);
    if (edge)
        dgemm_dmma_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(A, B, C, (int)M, (int)N, (int)K);
    else
        dgemm_dmma_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(A, B, C, (int)M, (int)N, (int)K);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // namespace dsb
