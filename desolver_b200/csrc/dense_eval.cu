// dense_eval.cu -- K5 evaluator: cubic Hermite dense output over the per-trajectory node store
// written by the fused ERK kernel.  Replaces DenseOutput.__call__ / find_interval
// (differential_system.py:205-227: first step whose END time >= t, clamped) and
// CubicHermiteInterp.__call__ (utilities/interpolation.py:41-61), one thread per trajectory.
#include "common.cuh"
#include <math_constants.h>

namespace dsb {

__global__ void dense_eval_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ node_count,
                                  long long n, int dim, int cap, const double *__restrict__ tq,
                                  long long tq_stride, double *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rec = 1 + 2 * dim;
    const double *base = nodes + i * (long long)cap * rec;
    int cnt = node_count[i];
    if (cnt > cap) cnt = cap;
    const double t = tq[i * tq_stride];
    if (cnt < 2) {   // no step recorded: return the only node (or NaN if none)
        for (int d = 0; d < dim; ++d) out[(long long)d * n + i] = cnt == 1 ? base[1 + d] : CUDART_NAN;
        return;
    }
    // steps k = 0..cnt-2 span nodes k -> k+1; find the first k whose end time t_{k+1} >= t
    int lo = 0, hi = cnt - 2;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (base[(long long)(mid + 1) * rec] >= t) hi = mid; else lo = mid + 1;
    }
    const double *n0 = base + (long long)lo * rec, *n1 = n0 + rec;
    const double t0 = n0[0], t1 = n1[0];
    const double trange = __dsub_rn(t1, t0);
    const double x = __ddiv_rn(__dsub_rn(t, t0), trange);
    if (x == 0.0 || x == 1.0) {
        const double *src = x == 0.0 ? n0 : n1;
        for (int d = 0; d < dim; ++d) out[(long long)d * n + i] = src[1 + d];
        return;
    }
    // same operation order as the reference, one rounding per operator
    const double x2 = __dmul_rn(x, x), x3 = __dmul_rn(x2, x);
    const double h00 = __dadd_rn(__dsub_rn(__dmul_rn(2.0, x3), __dmul_rn(3.0, x2)), 1.0);
    const double h10 = __dadd_rn(__dsub_rn(x3, __dmul_rn(2.0, x2)), x);
    const double h01 = __dadd_rn(__dmul_rn(-2.0, x3), __dmul_rn(3.0, x2));
    const double h11 = __dsub_rn(x3, x2);
    const double h10t = __dmul_rn(h10, trange), h11t = __dmul_rn(h11, trange);
    for (int d = 0; d < dim; ++d) {
        double v = __dmul_rn(h00, n0[1 + d]);
        v = __dadd_rn(v, __dmul_rn(h10t, n0[1 + dim + d]));
        v = __dadd_rn(v, __dmul_rn(h01, n1[1 + d]));
        v = __dadd_rn(v, __dmul_rn(h11t, n1[1 + dim + d]));
        out[(long long)d * n + i] = v;
    }
}

// Appends one node (t, y, f) to the store of every trajectory whose flag word has `mask` set (all trajectories
// when flags == nullptr): the dense-output writer of the stage-level path (the fused kernel writes its nodes
// itself).  One thread per trajectory; reads are coalesced over the ensemble axis.
__global__ void dense_append_kernel(double *__restrict__ nodes, int32_t *__restrict__ node_count, int cap, long long n,
                                    int dim, const double *__restrict__ t, const double *__restrict__ y,
                                    const double *__restrict__ f, const int32_t *__restrict__ flags, int mask)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags && !(flags[i] & mask)) return;
    const int c = node_count[i];
    if (c < cap) {
        double *rec = nodes + (i * (long long)cap + c) * (1 + 2 * dim);
        rec[0] = t[i];
        for (int d = 0; d < dim; ++d) {
            rec[1 + d] = y[(long long)d * n + i];
            rec[1 + dim + d] = f[(long long)d * n + i];
        }
    }
    node_count[i] = c + 1;
}

int launch_dense_append(double *nodes, int32_t *node_count, int32_t cap, int64_t n, int32_t dim, const double *t,
                        const double *y, const double *f, const int32_t *flags, int32_t mask, cudaStream_t stream)
{
    if (n <= 0) return DSB_OK;
    dense_append_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(nodes, node_count, cap, n, dim, t, y, f, flags, mask);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int launch_dense_eval(const double *nodes, const int32_t *node_count, int64_t n, int32_t dim, int32_t cap,
                      const double *tq, int64_t tq_stride, double *out, cudaStream_t stream)
{
    if (n <= 0) return DSB_OK;
    const int threads = 128;
    const long long blocks = (n + threads - 1) / threads;
    dense_eval_kernel<<<(unsigned)blocks, threads, 0, stream>>>(nodes, node_count, n, dim, cap, tq, tq_stride, out);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // namespace dsb
