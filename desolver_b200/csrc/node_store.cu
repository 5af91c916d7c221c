// node_store.cu -- dense output at ensemble scale: the paged, structure-of-arrays node store and its readers.
//
// Replaces the reference's per-step Python objects -- TableauIntegrator.dense_output (integrators/
// integrator_types.py:109-119: one CubicHermiteInterp per accepted step), DenseOutput.add_interpolant / find_interval /
// __call__ / grad (differential_system.py:170-290) and CubicHermiteInterp.__call__ / grad (utilities/
// interpolation.py:41-78) -- and also carries the reference's per-step history of a single system (`a.t`, `a.y`,
// differential_system.py:1015-1018).
//
// Layout (dsb_node_store, include/desolver_b200.h).  A node is (trajectory, ordinal k, t, y[dim], f[dim] = rhs(t, y)).
// The pool is a LOG of entries in groups of `width` entries; a group holds F = 2 + 2*dim rows of `width` doubles
// (row 0 = header {int32 trajectory, int32 k}, row 1 = t, rows 2.. = y, then f), so the entry e lives in slot e % width
// of group e / width.  Writers:
//   * the fused kernels (erk_fused.cuh, DENSE): width = 32.  A warp claims `page_entries` consecutive entries at a time
//     (one atomicAdd per page) and packs the nodes its lanes accept in one iteration into consecutive slots: every
//     field row is written as one or two contiguous segments, whatever trajectories the lanes currently hold and
//     however many nodes each of them has -- the ragged per-trajectory structure costs nothing while writing.
//   * dsb_node_append_rows (stage-level path, lock-step attempts): width = n, slot = trajectory, one row-group per call.
// Readers go through an index built after the run: offsets[i] = exclusive prefix sum of node_count (CSR), and
// loc[offsets[i] + k] = entry of node k of trajectory i (dsb_node_offsets, dsb_node_locate).
#include "common.cuh"
#include <math_constants.h>

namespace dsb {

__device__ __forceinline__ const double *node_field(const dsb_node_store &st, long long e, int field)
{
    const long long g = e / st.width, s = e - g * st.width;
    return st.pool + (g * (2 + 2 * st.dim) + field) * st.width + s;
}

// ---- index: exclusive scan of node_count -> offsets[n + 1] ---------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long *total)
{
    __shared__ long long s_warp[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        long long x = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0, xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long u = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += u;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = xi - x;          // exclusive warp offsets
        if (lane == 31 && total) *total = xi;
    }
    __syncthreads();
    const long long res = inc - v + s_warp[w];
    __syncthreads();
    return res;
}

// pass 1: tile sums; pass 2 (one block): exclusive scan of the tile sums, total -> offsets[n]; pass 3: tile-local scan
__global__ void scan_tile_sums_kernel(const int32_t *__restrict__ cnt, long long n, long long *__restrict__ tile_sum)
{
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
        if (i < n) v += cnt[i];
    }
    __shared__ long long s_total;
    block_exclusive_scan(v, &s_total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s_total;
}

__global__ void scan_tiles_kernel(long long *__restrict__ tile_sum, long long tiles, long long *__restrict__ offsets, long long n)
{
    __shared__ long long s_total, s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long b = 0; b < tiles; b += SCAN_THREADS) {
        const long long i = b + threadIdx.x;
        const long long v = i < tiles ? tile_sum[i] : 0;
        const long long ex = block_exclusive_scan(v, &s_total);
        if (i < tiles) tile_sum[i] = ex + s_carry;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = s_carry;
}

__global__ void scan_finish_kernel(const int32_t *__restrict__ cnt, long long n, const long long *__restrict__ tile_sum,
                                   long long *__restrict__ offsets)
{
    // thread t owns SCAN_ITEMS consecutive elements, so that one block-wide scan of the per-thread sums suffices
    const long long first = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    long long v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = first + k < n ? cnt[first + k] : 0;
        sum += v[k];
    }
    long long run = block_exclusive_scan(sum, nullptr) + tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (first + k < n) offsets[first + k] = run;
        run += v[k];
    }
}

// ---- index: loc[offsets[i] + k] = e for every valid entry ------------------------------------------------------
__global__ void node_locate_kernel(dsb_node_store st, long long claimed, long long n, const long long *__restrict__ offsets,
                                   long long *__restrict__ loc, unsigned long long *__restrict__ bad)
{
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < claimed; e += (long long)gridDim.x * blockDim.x) {
        if (st.page_fill) {
            const long long p = e / st.page_entries;
            if (e - p * st.page_entries >= st.page_fill[p]) continue;
        }
        const int2 hdr = *reinterpret_cast<const int2 *>(node_field(st, e, 0));
        if (hdr.x < 0) continue;
        const long long i = hdr.x;
        if (i >= n || hdr.y < 0 || offsets[i] + hdr.y >= offsets[i + 1]) { atomicAdd(bad, 1ull); continue; }
        loc[offsets[i] + hdr.y] = e;
    }
}

// ---- row writer (stage-level path) ------------------------------------------------------------------------------
// claim: one thread reserves the next row-group; cursor[2] = its first entry (-1: pool full, cursor[1] += n)
__global__ void node_claim_row_kernel(dsb_node_store st, long long n)
{
    unsigned long long *c = (unsigned long long *)st.cursor;
    const unsigned long long e0 = c[0];
    if ((long long)(e0 + n) <= st.capacity) { c[0] = e0 + n; c[2] = e0; }
    else { c[1] += (unsigned long long)n; c[2] = ~0ull; }
}

__global__ void node_append_rows_kernel(dsb_node_store st, long long n, const double *__restrict__ t,
                                        const double *__restrict__ y, const double *__restrict__ f,
                                        const int32_t *__restrict__ flags, int mask)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool want = !flags || (flags[i] & mask);
    const unsigned long long e0 = st.cursor[2];
    const bool room = e0 != ~0ull;
    if (blockIdx.y == 0) {
        int k = 0;
        if (want) { k = st.node_count[i]; st.node_count[i] = k + 1; }
        if (room) {
            double *hdr = const_cast<double *>(node_field(st, (long long)e0 + i, 0));
            *reinterpret_cast<int2 *>(hdr) = make_int2(want ? (int)i : -1, k);
            if (want) const_cast<double *>(node_field(st, (long long)e0 + i, 1))[0] = t[i];
        }
    }
    if (!want || !room) return;
    const int dim = st.dim;
    for (int d = blockIdx.y; d < dim; d += gridDim.y) {
        const_cast<double *>(node_field(st, (long long)e0 + i, 2 + d))[0] = y[(long long)d * n + i];
        if (f) const_cast<double *>(node_field(st, (long long)e0 + i, 2 + dim + d))[0] = f[(long long)d * n + i];
    }
}

// ---- readers ------------------------------------------------------------------------------------------------------
// Hermite basis in the reference's operation order, one rounding per operator (utilities/interpolation.py:41-78).
struct HermiteW { double w00, w10, w01, w11; };      // weights of p0, m0, p1, m1 (w10, w11 already times trange)

__device__ __forceinline__ HermiteW hermite_value_weights(double x, double trange)
{
    const double x2 = __dmul_rn(x, x), x3 = __dmul_rn(x2, x);
    HermiteW w;
    w.w00 = __dadd_rn(__dsub_rn(__dmul_rn(2.0, x3), __dmul_rn(3.0, x2)), 1.0);
    w.w10 = __dmul_rn(__dadd_rn(__dsub_rn(x3, __dmul_rn(2.0, x2)), x), trange);
    w.w01 = __dadd_rn(__dmul_rn(-2.0, x3), __dmul_rn(3.0, x2));
    w.w11 = __dmul_rn(__dsub_rn(x3, x2), trange);
    return w;
}

// grad(): t2 = 2 (t - t0)/trange (1/trange), t3 = 3 (t - t0)/trange (t - t0)/trange (1/trange)   (:68-69)
__device__ __forceinline__ HermiteW hermite_grad_weights(double t, double t0, double trange)
{
    const double inv = __ddiv_rn(1.0, trange), dt = __dsub_rn(t, t0);
    const double t2 = __dmul_rn(__ddiv_rn(__dmul_rn(2.0, dt), trange), inv);
    const double t3 = __dmul_rn(__ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(3.0, dt), trange), dt), trange), inv);
    HermiteW w;
    w.w00 = __dsub_rn(__dmul_rn(2.0, t3), __dmul_rn(3.0, t2));
    w.w10 = __dmul_rn(__dadd_rn(__dsub_rn(t3, __dmul_rn(2.0, t2)), inv), trange);
    w.w01 = __dadd_rn(__dmul_rn(-2.0, t3), __dmul_rn(3.0, t2));
    w.w11 = __dmul_rn(__dsub_rn(t3, t2), trange);
    return w;
}

__device__ __forceinline__ double hermite_combine(const HermiteW &w, double p0, double m0, double p1, double m1)
{
    double v = __dmul_rn(w.w00, p0);
    v = __dadd_rn(v, __dmul_rn(w.w10, m0));
    v = __dadd_rn(v, __dmul_rn(w.w01, p1));
    return __dadd_rn(v, __dmul_rn(w.w11, m1));
}

// out[q][d][i] = interpolant (derivative = 0) or its time derivative (1) of trajectory i at tq[i*sq_traj + q*sq_query].
// Interval lookup like DenseOutput.find_interval: the first step whose END time is >= t (<= t for backward integration),
// clamped to the last step.
__global__ void node_eval_kernel(dsb_node_store st, const long long *__restrict__ offsets, const long long *__restrict__ loc,
                                 long long n, const double *__restrict__ tq, long long sq_traj, long long sq_query, int nq,
                                 double *__restrict__ out, int derivative)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dim = st.dim;
    const long long base = offsets[i];
    const long long cnt = offsets[i + 1] - base;
    for (int q = blockIdx.y; q < nq; q += gridDim.y) {
        double *o = out + ((long long)q * dim) * n + i;
        if (cnt < 2) {      // no step recorded: the only node (or NaN if none)
            const long long e = cnt == 1 ? loc[base] : 0;
            for (int d = 0; d < dim; ++d) o[(long long)d * n] = cnt == 1 ? *node_field(st, e, (derivative ? 2 + dim : 2) + d) : CUDART_NAN;
            continue;
        }
        const double t = tq[i * sq_traj + (long long)q * sq_query];
        const double t_first = *node_field(st, loc[base], 1), t_last = *node_field(st, loc[base + cnt - 1], 1);
        const bool fwd = t_last >= t_first;
        long long lo = 0, hi = cnt - 2;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const double te = *node_field(st, loc[base + mid + 1], 1);
            if (fwd ? te >= t : te <= t) hi = mid; else lo = mid + 1;
        }
        const long long e0 = loc[base + lo], e1 = loc[base + lo + 1];
        const double t0 = *node_field(st, e0, 1), t1 = *node_field(st, e1, 1);
        const double trange = __dsub_rn(t1, t0);
        const double x = __ddiv_rn(__dsub_rn(t, t0), trange);
        if (x == 0.0 || x == 1.0) {
            const long long e = x == 0.0 ? e0 : e1;
            for (int d = 0; d < dim; ++d) o[(long long)d * n] = *node_field(st, e, (derivative ? 2 + dim : 2) + d);
            continue;
        }
        const HermiteW w = derivative ? hermite_grad_weights(t, t0, trange) : hermite_value_weights(x, trange);
        for (int d = 0; d < dim; ++d)
            o[(long long)d * n] = hermite_combine(w, *node_field(st, e0, 2 + d), *node_field(st, e0, 2 + dim + d),
                                                  *node_field(st, e1, 2 + d), *node_field(st, e1, 2 + dim + d));
    }
}

// nodes k0 .. k0 + count - 1 of ONE trajectory: t_out[j], y_out[j][dim] -- the reference's per-step history rows
__global__ void node_gather_kernel(dsb_node_store st, const long long *__restrict__ offsets, const long long *__restrict__ loc,
                                   long long traj, long long k0, long long count, double *__restrict__ t_out,
                                   double *__restrict__ y_out, double *__restrict__ f_out)
{
    const int dim = st.dim;
    const long long total = count * (dim + 1);
    const long long base = offsets[traj] + k0;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long j = x / (dim + 1);
        const int d = (int)(x - j * (dim + 1));
        const long long e = loc[base + j];
        if (d == dim) t_out[j] = *node_field(st, e, 1);
        else {
            y_out[j * dim + d] = *node_field(st, e, 2 + d);
            if (f_out) f_out[j * dim + d] = *node_field(st, e, 2 + dim + d);
        }
    }
}

// One cubic Hermite step given explicitly (the integrator protocol's dense_output(), integrator_types.py:109-119):
// out[d][i] over p0, p1, m0, m1 [dim][n], t0, t1 [n * st] (st = 0: shared), tq [n * sq]
__global__ void hermite_eval_kernel(const double *__restrict__ t0, const double *__restrict__ t1, long long st,
                                    const double *__restrict__ p0, const double *__restrict__ p1,
                                    const double *__restrict__ m0, const double *__restrict__ m1, long long n, int dim,
                                    const double *__restrict__ tq, long long sq, double *__restrict__ out, int derivative)
{
    const long long total = n * dim;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long i = x % n;
        const double a = t0[i * st], b = t1[i * st], t = tq[i * sq];
        const double trange = __dsub_rn(b, a);
        const double xa = __ddiv_rn(__dsub_rn(t, a), trange);
        if (xa == 0.0) { out[x] = derivative ? m0[x] : p0[x]; continue; }
        if (xa == 1.0) { out[x] = derivative ? m1[x] : p1[x]; continue; }
        const HermiteW w = derivative ? hermite_grad_weights(t, a, trange) : hermite_value_weights(xa, trange);
        out[x] = hermite_combine(w, p0[x], m0[x], p1[x], m1[x]);
    }
}

static int grid_for_total(long long total, int threads)
{
    long long b = (total + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b));
}

static int check_store(const dsb_node_store *st, const char *who)
{
    if (!st || !st->pool || !st->cursor || !st->node_count || st->dim < 1 || st->width < 1 || st->capacity < st->width ||
        st->capacity % st->width != 0) {
        set_error("%s: bad node store (pool, cursor, node_count required; capacity a positive multiple of width)", who);
        return DSB_ERR_BAD_ARGUMENT;
    }
    return DSB_OK;
}

}  // namespace dsb

using namespace dsb;

extern "C" {

int dsb_node_offsets(const int32_t *node_count, int64_t n, int64_t *offsets, int64_t *scratch, void *stream)
{
    if (!node_count || !offsets || !scratch || n < 1) { set_error("dsb_node_offsets: bad argument"); return DSB_ERR_BAD_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(node_count, n, (long long *)scratch);
    scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>((long long *)scratch, tiles, (long long *)offsets, n);
    scan_finish_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(node_count, n, (const long long *)scratch, (long long *)offsets);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int64_t dsb_node_offsets_scratch(int64_t n) { return n < 1 ? 0 : (n + SCAN_TILE - 1) / SCAN_TILE; }

int dsb_node_locate(const dsb_node_store *st, int64_t claimed_entries, int64_t n, const int64_t *offsets, int64_t *loc,
                    uint64_t *bad, void *stream)
{
    int rc = check_store(st, "dsb_node_locate");
    if (rc != DSB_OK) return rc;
    if (!offsets || !loc || !bad || claimed_entries < 0 || claimed_entries > st->capacity || n < 1) {
        set_error("dsb_node_locate: bad argument");
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (claimed_entries == 0) return DSB_OK;
    node_locate_kernel<<<grid_for_total(claimed_entries, 256), 256, 0, (cudaStream_t)stream>>>(
        *st, claimed_entries, n, (const long long *)offsets, (long long *)loc, (unsigned long long *)bad);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_node_append_rows(const dsb_node_store *st, int64_t n, const double *t, const double *y, const double *f,
                         const int32_t *flags, int32_t mask, void *stream)
{
    int rc = check_store(st, "dsb_node_append_rows");
    if (rc != DSB_OK) return rc;
    if (!t || !y || n < 1 || st->width != n) {
        set_error("dsb_node_append_rows: t and y are required and the store's width must equal n (%lld)", (long long)n);
        return DSB_ERR_BAD_ARGUMENT;
    }
    cudaStream_t s = (cudaStream_t)stream;
    node_claim_row_kernel<<<1, 1, 0, s>>>(*st, n);
    const unsigned gx = (unsigned)((n + 127) / 128);
    long long gy = st->dim;
    const long long cap = (148LL * 16 + gx - 1) / gx;             // enough blocks to fill the machine, no more
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    node_append_rows_kernel<<<dim3(gx, (unsigned)gy), 128, 0, s>>>(*st, n, t, y, f, flags, mask);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_node_eval(const dsb_node_store *st, const int64_t *offsets, const int64_t *loc, int64_t n, const double *tq,
                  int64_t tq_stride_traj, int64_t tq_stride_query, int32_t nq, double *out, int32_t derivative, void *stream)
{
    int rc = check_store(st, "dsb_node_eval");
    if (rc != DSB_OK) return rc;
    if (!offsets || !loc || !tq || !out || n < 1 || nq < 1) { set_error("dsb_node_eval: bad argument"); return DSB_ERR_BAD_ARGUMENT; }
    const unsigned gx = (unsigned)((n + 127) / 128);
    long long gy = nq;
    const long long cap = (148LL * 32 + gx - 1) / gx;
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    node_eval_kernel<<<dim3(gx, (unsigned)gy), 128, 0, (cudaStream_t)stream>>>(*st, (const long long *)offsets, (const long long *)loc, n,
                                                                              tq, tq_stride_traj, tq_stride_query, nq, out, derivative);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_node_gather(const dsb_node_store *st, const int64_t *offsets, const int64_t *loc, int64_t trajectory, int64_t k0,
                    int64_t count, double *t_out, double *y_out, double *f_out, void *stream)
{
    int rc = check_store(st, "dsb_node_gather");
    if (rc != DSB_OK) return rc;
    if (!offsets || !loc || !t_out || !y_out || trajectory < 0 || k0 < 0 || count < 0) {
        set_error("dsb_node_gather: bad argument");
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (count == 0) return DSB_OK;
    node_gather_kernel<<<grid_for_total(count * (st->dim + 1), 256), 256, 0, (cudaStream_t)stream>>>(
        *st, (const long long *)offsets, (const long long *)loc, trajectory, k0, count, t_out, y_out, f_out);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_hermite_eval(const double *t0, const double *t1, int64_t t_stride, const double *p0, const double *p1, const double *m0,
                     const double *m1, int64_t n, int32_t dim, const double *tq, int64_t tq_stride, double *out,
                     int32_t derivative, void *stream)
{
    if (!t0 || !t1 || !p0 || !p1 || !m0 || !m1 || !tq || !out || n < 1 || dim < 1) {
        set_error("dsb_hermite_eval: bad argument");
        return DSB_ERR_BAD_ARGUMENT;
    }
    hermite_eval_kernel<<<grid_for_total(n * dim, 256), 256, 0, (cudaStream_t)stream>>>(t0, t1, t_stride, p0, p1, m0, m1, n, dim, tq,
                                                                                       tq_stride, out, derivative);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // extern "C"
