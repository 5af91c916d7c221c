// ozaki_gemm.cu -- K3b: the dense linear right-hand side  K = A @ Y_stage  as an FP64-accurate product on the INT8 tensor
// cores of sm_100a (tcgen05.mma kind::i8, accumulators in TMEM, operands moved by TMA).
//
// Same job as dgemm.cu (the `A @ y` of the reference's linear benchmark system, evaluated by numpy/OpenBLAS inside
// compute_step, integrators/components/runge_kutta_methods.py:49-54), different arithmetic unit.  Blackwell has no FP64
// tcgen05 kind; its FP64 tensor path is the warp-level DMMA of dgemm.cu (37 TFLOP/s).  The INT8 path is 120 times wider,
// and integer products are exact, so the FP64 product is assembled from exact pieces (error-free splitting, the scheme
// of Ozaki, Ogita, Oishi & Rump):
//
//   * every row of A (every column of B) is scaled by a power of two so that |x| < 1 and cut into S = 8 signed digits:
//     x = q0 2^-6 + q1 2^-13 + ... + q7 2^-55 (+ a remainder below 2^-56), |q| <= 64, by round-to-nearest of the running
//     remainder -- every step exact in FP64.  55 bits relative to the largest element of the row (column);
//   * the products of digit planes A_i B_j are INT8 GEMMs with INT32 accumulation: |sum_k q q'| <= 4096 * 2^12 = 2^24, and
//     all planes of one weight i + j = d are summed in the same accumulator (at most 8 of them: < 2^27) -- exact;
//   * the 36 plane products with d = i + j <= 7 are kept (the dropped ones weigh < 2^-56 relative to |row| |column| K);
//     the 8 accumulators are combined in FP64 by Horner from the lightest weight: one rounding per accumulator.
//   The result is not the sequence of roundings of a left-to-right FP64 dot product (nor of cuBLAS' DMMA order); its
//   error is bounded by ~2^-52 |row max| |column max| K^0.5-ish, measured in tools/ozaki_check.py against an exact product.
//
// Kernel (one CTA per SM, persistent over 128 x 64 output tiles walked in bands of row blocks, 256 threads):
//   warp 0   TMA producer of the A tiles (128 x 64 bytes, 64-byte swizzle), ring of 16 slots with one mbarrier pair each;
//   warp 3   TMA producer of the B tiles (64 x 64 bytes), two groups of 8 slots (one group per k-block parity);
//            position p of a k-block brings A_(7-p) and B_p;
//   warp 1   MMA issuer (one elected lane): step p multiplies A_(7-p) with B_0 .. B_p (two K = 32 instructions each) into
//            accumulator d = 7 - p + j (TMEM columns 64 d .. 64 d + 63: all 512 columns).  A_(7-p)[kk] is read from shared
//            memory once and kept in the tensor core's collector buffer while it meets the p + 1 B tiles (.collector::a::fill /
//            use / lastuse: 3.53 -> 2.62 ms of MMA time).  The A slot is committed back to its producer after its step, the
//            B group at the end of the k-block;
//   warp 2   owns the TMEM allocation;
//   warps 4-7 epilogue: tcgen05.ld of the 8 accumulators of 16 columns, Horner in FP64, scale, 128-byte stores per lane.
// Two flavours: ozaki_gemm_kernel (every CTA on its own, any M % 128 == 0) and ozaki_gemm_pair_kernel (M % 256 == 0: CTA
// pairs, tcgen05 cta_group::2 -- 256 x 64 tiles, each CTA loads its 128 rows of A and HALF of the B rows, the leader
// issues UTCIMMA.2CTA for both).  Measured (4096 x 8192 x 4096, B200): pairs **2.93 ms**, single CTAs 3.65 ms, against
// 7.78 ms for cuBLAS DGEMM and 8.02 ms for dgemm.cu; MMA alone 2.55 ms, TMA alone 2.33 ms, the floor of the pipe 2.2 ms.
// In the single-CTA kernel the MMA lane never waits for operands (0.03 % of its cycles, DSB_OZAKI_MODE=4) but the tensor
// pipe runs at 52 cycles per instruction against a floor of 32: the shared-memory port carries the B tile of every
// instruction, the A fills and the TMA writes; pairs halve the B reads and cut the TMA bytes by a sixth.
// Tried and dropped: clusters of two CTAs multicasting the A tiles (exact, no gain -- L2 is not the limit); in the pair
// kernel a remote mbarrier arrive per load (11.5 ms) and cluster-scope acquire waits (4.9 ms).
//
// The digit planes of A are produced once per matrix (dsb_ozaki_split_a), those of Y per product (dsb_ozaki_split_b,
// which also transposes to K-major: [plane][column][k]).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace dsb {
#ifdef DSB_STANDALONE
void set_error(const char *, ...) {}
#endif

namespace oz {
#ifndef DSB_OZAKI_COLLECTOR
#define DSB_OZAKI_COLLECTOR 1
#endif
#ifndef DSB_OZAKI_COLL_MIN_P
#define DSB_OZAKI_COLL_MIN_P 1   // steps with fewer than this many + 1 products do not use the collector
#endif
#ifndef DSB_OZAKI_SLICES
#define DSB_OZAKI_SLICES 8
#endif
constexpr int S = DSB_OZAKI_SLICES;
constexpr int BM = 128, BN = 64, BK = 64, UMMA_K = 32;
constexpr int A_TILE = BM * BK, B_TILE = BN * BK, UNIT = A_TILE + B_TILE;
constexpr int RING = 2 * S;                      // slots per operand: two k-blocks
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = RING * UNIT + 1024 /* alignment slack */ + 512 /* barriers */;
constexpr int TMEM_COLS = 512;
static_assert(S * BN <= TMEM_COLS, "accumulators must fit TMEM");
constexpr long long WAIT_LIMIT = 4000000000LL;   // cycles (~2 s): a protocol error traps instead of hanging the GPU

__device__ __forceinline__ uint32_t sptr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int what)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > WAIT_LIMIT) {
            printf("ozaki_gemm: barrier wait %d timed out (block %d thread %d)\n", what, (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// the same, adding the cycles spent waiting to `acc` (mode & 4: where does the pipeline stall)
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, int what, long long &acc)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    mbar_wait(bar, parity, what);
    acc += clock64() - t0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// ---- CTA pairs (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: the data lands in THIS CTA's shared memory, the bytes are counted on the barrier `bar_cluster`
// (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// completion of all MMAs issued so far, signalled on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, signed 8-bit operands, 32-bit integer accumulators.  COLL: what happens to the A operand
// in the tensor core's collector buffer -- 0 discard (default), 1 fill (read from shared memory and keep), 2 use (A is the
// operand already in the buffer: no shared-memory read), 3 last use.
template <int COLL>
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
#define DSB_OZ_MMA(MOD)                                                                                   \
    asm volatile(                                                                                         \
        "{\n"                                                                                             \
        ".reg .pred p;\n"                                                                                 \
        "setp.ne.b32 p, %4, 0;\n"                                                                         \
        "tcgen05.mma.cta_group::1.kind::i8" MOD " [%0], %1, %2, %3, p;\n"                                 \
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory")
    if constexpr (COLL == 1) DSB_OZ_MMA(".collector::a::fill");
    else if constexpr (COLL == 2) DSB_OZ_MMA(".collector::a::use");
    else if constexpr (COLL == 3) DSB_OZ_MMA(".collector::a::lastuse");
    else DSB_OZ_MMA("");
#undef DSB_OZ_MMA
}
// the same for a CTA pair: D is 256 x N (128 rows in each CTA's TMEM), A = this CTA's 128 rows, B = N / 2 rows from each CTA
template <int COLL>
__device__ __forceinline__ void mma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
#define DSB_OZ_MMA2(MOD)                                                                                  \
    asm volatile(                                                                                         \
        "{\n"                                                                                             \
        ".reg .pred p;\n"                                                                                 \
        "setp.ne.b32 p, %4, 0;\n"                                                                         \
        "tcgen05.mma.cta_group::2.kind::i8" MOD " [%0], %1, %2, %3, p;\n"                                 \
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory")
    if constexpr (COLL == 1) DSB_OZ_MMA2(".collector::a::fill");
    else if constexpr (COLL == 2) DSB_OZ_MMA2(".collector::a::use");
    else if constexpr (COLL == 3) DSB_OZ_MMA2(".collector::a::lastuse");
    else DSB_OZ_MMA2("");
#undef DSB_OZ_MMA2
}
// K-major operand tile whose rows are 64 bytes, 64-byte swizzle (TMA CU_TENSOR_MAP_SWIZZLE_64B): groups of 8 rows are
// 512 bytes apart (stride byte offset); the leading byte offset is not used by swizzled K-major layouts
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr)
{
    const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (512u >> 4) | (1u << 14) /* descriptor version of sm_100 */ | (4u << 29) /* SWIZZLE_64B */;
    return ((uint64_t)hi << 32) | lo;
}
// c_format S32 (2) | a, b INT8 (1) | K-major both | N >> 3 | M >> 4
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr int B_HALF = B_TILE / 2;               // a CTA of a pair holds half of the B rows

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace oz

// One output row of a tile (this lane's TMEM lane): the 8 accumulators of 16 columns at a time, Horner in FP64 from the
// lightest weight, the two power-of-two scales, 128-byte stores.
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_base, int q, int row, int n0, const double *__restrict__ scale_a,
                                              const double *__restrict__ scale_b, double *__restrict__ C, int M, int N,
                                              int *__restrict__ dbg_acc)
{
    using namespace oz;
    constexpr double W = 1.0 / 128.0;                                         // 2^-7 between neighbouring weights
    const double sa = scale_a[row] * (1.0 / 4096.0);                          // 2^-6 * 2^-6 of the two leading digits
    double *crow = C + (size_t)row * N + n0;
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
        int acc[S][16];
#pragma unroll
        for (int d = 0; d < S; ++d) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * BN + c * 16), acc[d]);
        tmem_ld_wait();
        if (dbg_acc != nullptr) {
#pragma unroll
            for (int d = 0; d < S; ++d)
#pragma unroll
                for (int e = 0; e < 16; ++e) dbg_acc[((size_t)d * M + row) * N + n0 + c * 16 + e] = acc[d][e];
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            double v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double s = (double)acc[S - 1][e + h];
#pragma unroll
                for (int d = S - 2; d >= 0; --d) s = fma(s, W, (double)acc[d][e + h]);
                v[h] = s * sa * __ldg(scale_b + n0 + c * 16 + e + h);
            }
            *reinterpret_cast<double2 *>(crow + c * 16 + e) = make_double2(v[0], v[1]);
        }
    }
}

__global__ void __launch_bounds__(oz::THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const double *__restrict__ scale_a, const double *__restrict__ scale_b, double *__restrict__ C, int M, int N,
                  int K, int *__restrict__ dbg_acc, int mode, int band)
{
    using namespace oz;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (sptr(smem_raw) + 1023u) & ~1023u;             // swizzled tiles: 1024-byte aligned
    const uint32_t smem_b = smem_base + RING * A_TILE;                        // A slots, then B slots
    const uint32_t bars = smem_b + RING * B_TILE;
    auto full_a = [&](int u) { return bars + 8u * u; };
    auto full_b = [&](int u) { return bars + 8u * (RING + u); };
    auto empty_a = [&](int u) { return bars + 8u * (2 * RING + u); };
    auto empty_b = [&](int g) { return bars + 8u * (3 * RING + g); };          // one per k-block parity: B tiles retire together
    const uint32_t tmem_full = bars + 8u * (3 * RING + 2), tmem_empty = tmem_full + 8u, tmem_slot = tmem_empty + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = M / BM, tiles = mb * (N / BN), kblocks = K / BK;
    // work item `it` of this CTA: row block m0, column block n0
    auto tile_at = [&](int it, int &m0, int &n0) {
        const int g = blockIdx.x + it * gridDim.x;
        if (g >= tiles) return false;
        // row blocks are walked in bands of `band` (m fastest inside a band): the CTAs running at the same time share
        // band * 4 MB of A planes and ~(SMs / band) * 2 MB of B planes, which stay in L2 from one wave to the next
        const int nbc = N / BN, per_band = band * nbc, b0 = (g / per_band) * band;
        const int rows = (mb - b0) < band ? (mb - b0) : band, r = g - (g / per_band) * per_band;
        m0 = (b0 + r % rows) * BM;
        n0 = (r / rows) * BN;
        return true;
    };

    if (warp == 1 && lane == 0) {
        for (int u = 0; u < RING; ++u) { mbar_init(full_a(u), 1); mbar_init(full_b(u), 1); mbar_init(empty_a(u), 1); }
        mbar_init(empty_b(0), 1);
        mbar_init(empty_b(1), 1);
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // Order inside a k-block: position p = 0 .. S-1 brings A_(S-1-p) and B_p; step p multiplies A_(S-1-p) with B_0 .. B_p.
    // The MMA warp can start on the first tiles that land, an A tile is dead after its own step (its slot goes back to the
    // producer at once), the B tiles of a k-block retire together at its end.
    if (warp == 0) {
      if (lane == 0 && !(mode & 1)) {
        // ===== TMA producer of the A tiles =====
        uint32_t n_unit = 0;
        int m0, n0;
        long long waited = 0;
        for (int it = 0; tile_at(it, m0, n0); ++it) {
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll 1
                for (int p = 0; p < S; ++p, ++n_unit) {
                    const int slot = n_unit % RING;
                    const uint32_t use = n_unit / RING;
                    if (use > 0) mbar_wait_timed(empty_a(slot), (use - 1) & 1, 100 + slot, waited);
                    mbar_expect_tx(full_a(slot), A_TILE);
                    tma_load_2d(smem_base + slot * A_TILE, &map_a, kb * BK, (S - 1 - p) * M + m0, full_a(slot));
                }
            }
        }
        if (mode & 4) reinterpret_cast<long long *>(dbg_acc)[blockIdx.x * 4 + 1] = waited;
      }
      __syncwarp();
    } else if (warp == 3) {
      if (lane == 0 && !(mode & 1)) {
        // ===== TMA producer of the B tiles =====
        uint32_t n_kb = 0;
        int m0, n0;
        long long waited = 0;
        for (int it = 0; tile_at(it, m0, n0); ++it) {
            for (int kb = 0; kb < kblocks; ++kb, ++n_kb) {
                const int grp = n_kb & 1;
                const uint32_t use = n_kb >> 1;
                if (use > 0) mbar_wait_timed(empty_b(grp), (use - 1) & 1, 150 + grp, waited);
#pragma unroll 1
                for (int p = 0; p < S; ++p) {
                    const int slot = grp * S + p;
                    mbar_expect_tx(full_b(slot), B_TILE);
                    tma_load_2d(smem_b + slot * B_TILE, &map_b, kb * BK, p * N + n0, full_b(slot));
                }
            }
        }
        if (mode & 4) reinterpret_cast<long long *>(dbg_acc)[blockIdx.x * 4 + 2] = waited;
      }
      __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (warp-uniform descriptors), one elected lane issues =====
        const uint32_t desc_hi = (512u >> 4) | (1u << 14) | (4u << 29);
        const uint32_t lo_a0 = ((smem_base >> 4) & 0x3FFFu) | (1u << 16), lo_b0 = ((smem_b >> 4) & 0x3FFFu) | (1u << 16);
        auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
        uint32_t n_kb = 0, n_tile = 0;
        int m0, n0;
        long long waited = 0, waited_epi = 0;
        const long long t_start = clock64();
        for (int it = 0; tile_at(it, m0, n0); ++it, ++n_tile) {
            if (n_tile > 0) {
                mbar_wait_timed(tmem_empty, (n_tile - 1) & 1, 300, waited_epi);   // the epilogue has drained the accumulators
                tc_fence_after();
            }
            for (int kb = 0; kb < kblocks; ++kb, ++n_kb) {
                const int base = (n_kb & 1) * S;
                const uint32_t parity = (n_kb >> 1) & 1;
                const uint32_t lo_a = lo_a0 + (uint32_t)base * (A_TILE / 16), lo_b = lo_b0 + (uint32_t)base * (B_TILE / 16);
                if (elect_one()) {
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        if (!(mode & 1)) {
                            mbar_wait_timed(full_a(base + p), parity, 200 + base + p, waited);
                            mbar_wait_timed(full_b(base + p), parity, 250 + base + p, waited);
                            tc_fence_after();
                        }
                        if (!(mode & 2)) {
#pragma unroll
                            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                // A_(S-1-p)[kk] stays in the collector while it meets B_0 .. B_p
#pragma unroll
                                for (int j = 0; j <= p; ++j) {
                                    const uint32_t d_tmem = tmem_base + (uint32_t)((S - 1 - p + j) * BN);
                                    const uint64_t ad = desc(lo_a + p * (A_TILE / 16) + kk * (UMMA_K / 16));
                                    const uint64_t bd = desc(lo_b + j * (B_TILE / 16) + kk * (UMMA_K / 16));
                                    // the first product into accumulator d = S-1-p+j of a tile is (kb 0, kk 0, j 0)
                                    const uint32_t accumulate = (j > 0 || kk > 0) ? 1u : (uint32_t)(kb > 0);
                                    if (!DSB_OZAKI_COLLECTOR || p == 0) mma_i8<0>(d_tmem, ad, bd, IDESC, accumulate);
                                    else if (j == 0) mma_i8<1>(d_tmem, ad, bd, IDESC, accumulate);
                                    else if (j == p) mma_i8<3>(d_tmem, ad, bd, IDESC, accumulate);
                                    else mma_i8<2>(d_tmem, ad, bd, IDESC, accumulate);
                                }
                            }
                        }
                        tc_commit(empty_a(base + p));                          // this A tile is dead once these complete
                    }
                    tc_commit(empty_b(n_kb & 1));                              // ... and so are all B tiles of the k-block
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(tmem_full);
            __syncwarp();
        }
        if ((mode & 4) && lane == 0) {
            long long *st = reinterpret_cast<long long *>(dbg_acc) + blockIdx.x * 4;
            st[0] = clock64() - t_start;
            st[3] = waited_epi;
        }
        // the elected lane need not be lane 0: every lane adds what it waited (the others waited nothing)
        if (mode & 4) atomicAdd(reinterpret_cast<unsigned long long *>(dbg_acc) + gridDim.x * 4 + blockIdx.x, (unsigned long long)waited);
    } else if (warp >= 4) {
        // ===== epilogue: warp q reads TMEM lanes 32 q .. 32 q + 31 =====
        const int q = warp - 4;
        uint32_t n_tile = 0;
        int m0, n0;
        for (int it = 0; tile_at(it, m0, n0); ++it, ++n_tile) {
            mbar_wait(tmem_full, n_tile & 1, 400);
            tc_fence_after();
            epilogue_tile(tmem_base, q, m0 + q * 32 + lane, n0, scale_a, scale_b, C, M, N, (mode & 4) ? nullptr : dbg_acc);
            tc_fence_before();
            mbar_arrive(tmem_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ---- the same product on CTA pairs (tcgen05 cta_group::2) ---------------------------------------------------------------
// Two CTAs on the SMs of one TPC share a 256 x 64 output tile: each holds 128 rows of every accumulator in its own TMEM,
// loads its own 128 rows of the A planes and HALF of the B planes' rows (32 of 64) -- the tensor cores of the pair read the
// B operand from both shared memories, so a CTA moves 10 KB instead of 12 KB per unit and re-reads half as much B.  Only
// the leader (cluster rank 0) issues MMAs; TMA loads of both CTAs count on the leader's `full` barriers, the leader's
// commits release the slots and publish the accumulators in both CTAs.
__global__ void __launch_bounds__(oz::THREADS, 1)
ozaki_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const double *__restrict__ scale_a, const double *__restrict__ scale_b, double *__restrict__ C, int M,
                       int N, int K, int *__restrict__ dbg_acc, int mode, int band)
{
    using namespace oz;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (sptr(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_b = smem_base + RING * A_TILE;
    const uint32_t bars = smem_b + RING * B_HALF;
    auto full_a = [&](int u) { return bars + 8u * u; };
    auto full_b = [&](int u) { return bars + 8u * (RING + u); };
    auto empty_a = [&](int u) { return bars + 8u * (2 * RING + u); };
    auto empty_b = [&](int g) { return bars + 8u * (3 * RING + g); };
    const uint32_t tmem_full = bars + 8u * (3 * RING + 2), tmem_empty = tmem_full + 8u, tmem_slot = tmem_empty + 8u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int mb = M / (2 * BM), items = mb * (N / BN), kblocks = K / BK;
    // work item `it` of this pair: 256-row block (this CTA: its 128 rows m0), column block n0
    auto tile_at = [&](int it, int &m0, int &n0) {
        const int g = pair_id + it * n_pairs;
        if (g >= items) return false;
        const int nbc = N / BN, per_band = band * nbc, b0 = (g / per_band) * band;
        const int rows = (mb - b0) < band ? (mb - b0) : band, r = g - (g / per_band) * per_band;
        m0 = (b0 + r % rows) * (2 * BM) + (int)rank * BM;
        n0 = (r / rows) * BN;
        return true;
    };

    if (warp == 1 && lane == 0) {
        // `full` (used in the leader only): one arrival, the leader's, carrying the byte count of BOTH CTAs' loads -- the
        // peer's loads just complete their bytes there (a remote arrive per unit costs ~0.5 us: 11.5 ms instead of 3.6)
        for (int u = 0; u < RING; ++u) { mbar_init(full_a(u), 1); mbar_init(full_b(u), 1); mbar_init(empty_a(u), 1); }
        mbar_init(empty_b(0), 1);
        mbar_init(empty_b(1), 1);
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 256);                                            // the epilogue threads of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
      if (lane == 0 && !(mode & 1)) {
        // ===== TMA producer of this CTA's A tiles =====
        uint32_t n_unit = 0;
        const uint32_t lbar0 = map_to_cta(full_a(0), 0);
        int m0, n0;
        for (int it = 0; tile_at(it, m0, n0); ++it) {
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll 1
                for (int p = 0; p < S; ++p, ++n_unit) {
                    const int slot = n_unit % RING;
                    const uint32_t use = n_unit / RING;
                    if (use > 0) mbar_wait(empty_a(slot), (use - 1) & 1, 100 + slot);
                    const uint32_t lbar = lbar0 + 8u * slot;
                    if (leader) mbar_expect_tx(full_a(slot), 2 * A_TILE);
                    tma_load_2d_pair(smem_base + slot * A_TILE, &map_a, kb * BK, (S - 1 - p) * M + m0, lbar);
                }
            }
        }
      }
      __syncwarp();
    } else if (warp == 3) {
      if (lane == 0 && !(mode & 1)) {
        // ===== TMA producer of this CTA's half of the B tiles =====
        uint32_t n_kb = 0;
        const uint32_t lbar0 = map_to_cta(full_b(0), 0);
        int m0, n0;
        for (int it = 0; tile_at(it, m0, n0); ++it) {
            for (int kb = 0; kb < kblocks; ++kb, ++n_kb) {
                const int grp = n_kb & 1;
                const uint32_t use = n_kb >> 1;
                if (use > 0) mbar_wait(empty_b(grp), (use - 1) & 1, 150 + grp);
#pragma unroll 1
                for (int p = 0; p < S; ++p) {
                    const int slot = grp * S + p;
                    const uint32_t lbar = lbar0 + 8u * slot;
                    if (leader) mbar_expect_tx(full_b(slot), 2 * B_HALF);
                    tma_load_2d_pair(smem_b + slot * B_HALF, &map_b, kb * BK, p * N + n0 + (int)rank * (BN / 2), lbar);
                }
            }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (leader) {
        // ===== MMA issuer of the pair =====
        const uint32_t desc_hi = (512u >> 4) | (1u << 14) | (4u << 29);
        const uint32_t lo_a0 = ((smem_base >> 4) & 0x3FFFu) | (1u << 16), lo_b0 = ((smem_b >> 4) & 0x3FFFu) | (1u << 16);
        auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
        uint32_t n_kb = 0, n_tile = 0;
        int m0, n0;
        for (int it = 0; tile_at(it, m0, n0); ++it, ++n_tile) {
            if (n_tile > 0) {
                mbar_wait(tmem_empty, (n_tile - 1) & 1, 300);          // both epilogues have drained the accumulators
                tc_fence_after();
            }
            for (int kb = 0; kb < kblocks; ++kb, ++n_kb) {
                const int base = (n_kb & 1) * S;
                const uint32_t parity = (n_kb >> 1) & 1;
                const uint32_t lo_a = lo_a0 + (uint32_t)base * (A_TILE / 16), lo_b = lo_b0 + (uint32_t)base * (B_HALF / 16);
                if (elect_one()) {
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        if (!(mode & 1)) {
                            mbar_wait(full_a(base + p), parity, 200 + base + p);
                            mbar_wait(full_b(base + p), parity, 250 + base + p);
                            tc_fence_after();
                        }
                        if (!(mode & 2)) {
#pragma unroll
                            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
#pragma unroll
                                for (int j = 0; j <= p; ++j) {
                                    const uint32_t d_tmem = tmem_base + (uint32_t)((S - 1 - p + j) * BN);
                                    const uint64_t ad = desc(lo_a + p * (A_TILE / 16) + kk * (UMMA_K / 16));
                                    const uint64_t bd = desc(lo_b + j * (B_HALF / 16) + kk * (UMMA_K / 16));
                                    const uint32_t accumulate = (j > 0 || kk > 0) ? 1u : (uint32_t)(kb > 0);
                                    if (!DSB_OZAKI_COLLECTOR || p < DSB_OZAKI_COLL_MIN_P) mma_i8_pair<0>(d_tmem, ad, bd, IDESC_PAIR, accumulate);
                                    else if (j == 0) mma_i8_pair<1>(d_tmem, ad, bd, IDESC_PAIR, accumulate);
                                    else if (j == p) mma_i8_pair<3>(d_tmem, ad, bd, IDESC_PAIR, accumulate);
                                    else mma_i8_pair<2>(d_tmem, ad, bd, IDESC_PAIR, accumulate);
                                }
                            }
                        }
                        tc_commit_pair(empty_a(base + p));
                    }
                    tc_commit_pair(empty_b(n_kb & 1));
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit_pair(tmem_full);
            __syncwarp();
        }
      }
      __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue of this CTA's 128 rows =====
        const int q = warp - 4;
        uint32_t n_tile = 0;
        const uint32_t lempty = map_to_cta(tmem_empty, 0);
        int m0, n0;
        for (int it = 0; tile_at(it, m0, n0); ++it, ++n_tile) {
            mbar_wait(tmem_full, n_tile & 1, 400);
            tc_fence_after();
            epilogue_tile(tmem_base, q, m0 + q * 32 + lane, n0, scale_a, scale_b, C, M, N, (mode & 4) ? nullptr : dbg_acc);
            tc_fence_before();
            mbar_arrive_cluster(lempty);
        }
    }
    tc_fence_before();
    cluster_sync();                                                            // nobody leaves while its peer can still signal it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ---- digit planes ------------------------------------------------------------------------------------------------------
namespace oz {
// power of two p with max < p (max = f 2^e, 0.5 <= f < 1  ->  p = 2^e); 1 for an all-zero (or non-finite) vector
__device__ __forceinline__ double pow2_above(double mx)
{
    if (!(mx > 0.0) || !(mx < 1.7e308)) return 1.0;
    int e;
    frexp(mx, &e);
    return ldexp(1.0, e);
}
// the S signed digits of x (|x| < 1): q0 = rint(x 2^6), then 7 bits at a time of the exact remainder
__device__ __forceinline__ void digits(double x, signed char (&q)[S])
{
    double r = x * 64.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double n = rint(r);
        q[s] = (signed char)(int)n;
        r = (r - n) * 128.0;
    }
}
}  // namespace oz

// A[M][K] row-major -> planes[S][M][K] int8, scale[M]: one CTA per row, 16 consecutive elements per thread and pass
__global__ void __launch_bounds__(256) ozaki_split_rows_kernel(const double *__restrict__ A, int K, size_t plane_stride,
                                                               signed char *__restrict__ planes, double *__restrict__ scale)
{
    using namespace oz;
    const int row = blockIdx.x;
    const double *a = A + (size_t)row * K;
    __shared__ double red[8];
    double mx = 0.0;
    for (int k = threadIdx.x * 2; k < K; k += 512) {
        const double2 v = *reinterpret_cast<const double2 *>(a + k);
        mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
    const double p = pow2_above(mx), inv = 1.0 / p;
    if (threadIdx.x == 0) scale[row] = p;
    for (int k0 = threadIdx.x * 16; k0 < K; k0 += 256 * 16) {
        signed char out[S][16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(a + k0 + e);
            signed char q[S];
            digits(v.x * inv, q);
#pragma unroll
            for (int s = 0; s < S; ++s) out[s][e] = q[s];
            digits(v.y * inv, q);
#pragma unroll
            for (int s = 0; s < S; ++s) out[s][e + 1] = q[s];
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4 *>(planes + s * plane_stride + (size_t)row * K + k0) = *reinterpret_cast<const uint4 *>(out[s]);
    }
}

// column maxima of B[K][N] (N contiguous): bit pattern of |b| ordered like an unsigned integer, atomicMax per column
__global__ void __launch_bounds__(256) ozaki_colmax_kernel(const double *__restrict__ B, int K, int N, int rows_per_block,
                                                           unsigned long long *__restrict__ colmax_bits)
{
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const int k0 = blockIdx.y * rows_per_block, k1 = min(K, k0 + rows_per_block);
    double mx = 0.0;
    for (int k = k0; k < k1; ++k) mx = fmax(mx, fabs(B[(size_t)k * N + n]));
    atomicMax(colmax_bits + n, (unsigned long long)__double_as_longlong(mx));
}

// B[K][N] -> planes[S][N][K] int8 (transposed: k contiguous), scale[N].  A CTA takes 32 columns x 128 rows: lane = column
// (256-byte coalesced row segments), each thread packs the digits of 4 consecutive rows into one 32-bit word per plane,
// the words go through shared memory (rows padded to 132 bytes: conflict-free both ways) and leave as 128-byte runs
// along k, one (plane, column) per warp store.
__global__ void __launch_bounds__(256) ozaki_split_cols_kernel(const double *__restrict__ B, int K, int N, size_t plane_stride,
                                                               const unsigned long long *__restrict__ colmax_bits,
                                                               signed char *__restrict__ planes, double *__restrict__ scale)
{
    using namespace oz;
    constexpr int COLS = 32, ROWS = 128, LD = ROWS + 4;
    __shared__ __align__(16) unsigned char tile[S][COLS][LD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = blockIdx.x * COLS, k0 = blockIdx.y * ROWS, n = n0 + lane;
    if (n < N) {
        const double p = pow2_above(__longlong_as_double((long long)colmax_bits[n])), inv = 1.0 / p;
        if (blockIdx.y == 0 && warp == 0) scale[n] = p;
        double v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[r][e] = B[(size_t)(k0 + 4 * (warp + 8 * r) + e) * N + n];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t word[S];
#pragma unroll
            for (int s = 0; s < S; ++s) word[s] = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                signed char q[S];
                digits(v[r][e] * inv, q);
#pragma unroll
                for (int s = 0; s < S; ++s) word[s] |= (uint32_t)(unsigned char)q[s] << (8 * e);
            }
#pragma unroll
            for (int s = 0; s < S; ++s) *reinterpret_cast<uint32_t *>(&tile[s][lane][4 * (warp + 8 * r)]) = word[s];
        }
    }
    __syncthreads();
    // warp w writes (plane, column) pairs w, w + 8, ...: 128 contiguous bytes each
#pragma unroll 4
    for (int pair = warp; pair < S * COLS; pair += 8) {
        const int s = pair / COLS, c = pair % COLS;
        if (n0 + c < N)
            *reinterpret_cast<uint32_t *>(planes + s * plane_stride + (size_t)(n0 + c) * K + k0 + 4 * lane) =
                *reinterpret_cast<const uint32_t *>(&tile[s][c][4 * lane]);
    }
}

namespace oz {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// Tuning / diagnosis knobs, read once: DSB_OZAKI_MODE (bit 0: no loads, bit 1: no MMAs, bit 2: wait-cycle statistics into
// dbg_acc -- tools/ozaki_check.py, tools/ozaki_stalls.py), DSB_OZAKI_BAND (row blocks per band), DSB_OZAKI_NO_PAIR.
struct Knobs {
    int mode = 0, band = 0;
    bool no_pair = false;
    Knobs()
    {
        if (const char *e = getenv("DSB_OZAKI_MODE")) mode = atoi(e);
        if (const char *e = getenv("DSB_OZAKI_BAND")) band = atoi(e);
        no_pair = getenv("DSB_OZAKI_NO_PAIR") != nullptr;
    }
};
static const Knobs &knobs()
{
    static const Knobs k;
    return k;
}
// planes[S * rows][K] int8, box of `box_rows` rows x 64 bytes, 64-byte swizzle
static int make_map(CUtensorMap *map, const void *planes, long long rows, long long K, int box_rows)
{
    EncodeTiledFn fn = encode_tiled();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DSB_ERR_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)K};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(planes), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DSB_ERR_CUDA; }
    return DSB_OK;
}
}  // namespace oz
}  // namespace dsb

using namespace dsb;

extern "C" {

DSB_API int dsb_ozaki_slices(void) { return oz::S; }

DSB_API int dsb_ozaki_split_a(const double *A, int64_t M, int64_t K, signed char *planes, double *scale, void *stream)
{
    if (A == nullptr || planes == nullptr || scale == nullptr || M <= 0 || K <= 0 || K % 16 != 0) {
        set_error("dsb_ozaki_split_a: bad arguments (K must be a multiple of 16)");
        return DSB_ERR_BAD_ARGUMENT;
    }
    ozaki_split_rows_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(A, (int)K, (size_t)M * K, planes, scale);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

// workspace: N * 8 bytes (column maxima)
DSB_API int dsb_ozaki_split_b(const double *B, int64_t K, int64_t N, signed char *planes, double *scale, void *workspace,
                              void *stream)
{
    if (B == nullptr || planes == nullptr || scale == nullptr || workspace == nullptr || N <= 0 || K <= 0 || K % 128 != 0) {
        set_error("dsb_ozaki_split_b: bad arguments (K must be a multiple of 128)");
        return DSB_ERR_BAD_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DSB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, (size_t)N * 8, st));
    const int rows_per_block = 128;
    dim3 g1((unsigned)((N + 255) / 256), (unsigned)((K + rows_per_block - 1) / rows_per_block));
    ozaki_colmax_kernel<<<g1, 256, 0, st>>>(B, (int)K, (int)N, rows_per_block, (unsigned long long *)workspace);
    dim3 g2((unsigned)((N + 31) / 32), (unsigned)(K / 128));
    ozaki_split_cols_kernel<<<g2, 256, 0, st>>>(B, (int)K, (int)N, (size_t)N * K, (const unsigned long long *)workspace, planes, scale);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

// C[M][N] = A[M][K] B[K][N] from the digit planes of A ([S][M][K]) and of B ([S][N][K]); dbg_acc (optional, [S][M][N] int32)
// receives the raw accumulators
DSB_API int dsb_ozaki_gemm(const signed char *a_planes, const double *a_scale, const signed char *b_planes, const double *b_scale,
                           double *C, int64_t M, int64_t N, int64_t K, int sms, int32_t *dbg_acc, void *stream)
{
    using namespace oz;
    if (M <= 0 || N <= 0 || K <= 0 || M % BM != 0 || N % BN != 0 || K % BK != 0 || K > 4096 * 8) {
        set_error("dsb_ozaki_gemm: needs M %% 128 == 0, N %% 64 == 0, K %% 64 == 0 and K <= 32768 (INT32 accumulators)");
        return DSB_ERR_BAD_ARGUMENT;
    }
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, a_planes, (long long)S * M, K, BM);
    if (rc != DSB_OK) return rc;
    rc = make_map(&map_b, b_planes, (long long)S * N, K, BN);
    if (rc != DSB_OK) return rc;
    const int mode = knobs().mode;
    const int band = knobs().band > 0 ? knobs().band : 16;                    // single CTAs: bands of 16 x 128 rows
    if ((M % (2 * BM) == 0) && sms >= 2 && !knobs().no_pair) {
        // CTA pairs, as many as the device can hold at once (a pair needs both SMs of a TPC: on a part with odd TPCs fewer
        // than sms / 2 fit, and a persistent grid larger than that would run in two waves); with too few pairs, or if the
        // cluster launch is refused, the single-CTA kernel below takes over
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // per device, once: the shared-memory opt-in and how many pairs fit (both are host-side driver queries that must not
        // sit in front of every product)
        static int resident_of[64];
        static bool known[64];
        int device = 0;
        DSB_CUDA_CHECK(cudaGetDevice(&device));
        const int slot = device & 63;
        if (!known[slot]) {
            DSB_CUDA_CHECK(cudaFuncSetAttribute(ozaki_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            DSB_CUDA_CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            int r = 0;
            if (cudaOccupancyMaxActiveClusters(&r, ozaki_gemm_pair_kernel, &cfg) != cudaSuccess) {
                cudaGetLastError();
                r = 0;
            }
            resident_of[slot] = r;
            known[slot] = true;
        }
        const int resident = resident_of[slot];
        const int want = sms / 2 < resident ? sms / 2 : resident;
        if (want * 10 >= (sms / 2) * 8) {
            CUtensorMap map_bh;
            rc = make_map(&map_bh, b_planes, (long long)S * N, K, BN / 2);
            if (rc != DSB_OK) return rc;
            const long long items = (M / (2 * BM)) * (N / BN);
            cfg.gridDim = dim3((unsigned)(2 * (items < want ? items : want)));
            const int pband = knobs().band > 0 ? knobs().band : 4;             // pairs: bands of 4 x 256 rows
            if (cudaLaunchKernelEx(&cfg, ozaki_gemm_pair_kernel, map_a, map_bh, a_scale, b_scale, C, (int)M, (int)N, (int)K, dbg_acc,
                                   mode, pband) == cudaSuccess)
                return DSB_OK;
            cudaGetLastError();
        }
    }
    DSB_CUDA_CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const long long tiles = (M / BM) * (N / BN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    ozaki_gemm_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_b, a_scale, b_scale, C, (int)M, (int)N, (int)K,
                                                                           dbg_acc, mode, band);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}
}
