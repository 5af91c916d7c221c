// host_pipeline.cu -- dsb_erk_run_host: an ensemble that lives in HOST memory goes through ONE persistent
// launch of the fused kernel while its own uploads and downloads are in flight (include/desolver_b200.h).
//
//   stream_up     memcpy2D(chunk 0) . write flags[0]=len0 . memcpy2D(chunk 1) . write flags[0]=len0+len1 ...
//   stream        memset(flags) ........ erk_fused_kernel (polls flags[0] before loading, counts flags[1+c]) ........ wait(down)
//   stream_down   wait flags[1] == len0 . D2H(chunk 0) . wait flags[2] == len1 . D2H(chunk 1) ...
//
// The stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) are driver-API calls; they are
// resolved through cudaGetDriverEntryPoint, so the library still links against cudart only.
#include <cuda.h>

#include "plan.h"

namespace dsb {

typedef CUresult (*write_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*wait_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static write_value_fn g_write_value = nullptr;
static wait_value_fn g_wait_value = nullptr;
static int g_memops_state = 0;      // 0 unknown, 1 available, -1 unavailable

static bool resolve_memops(int device)
{
    if (g_memops_state != 0) return g_memops_state > 0;
    g_memops_state = -1;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult rw, rq;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &rw) != cudaSuccess ||
        cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &rq) != cudaSuccess ||
        rw != cudaDriverEntryPointSuccess || rq != cudaDriverEntryPointSuccess || !w || !q) {
        cudaGetLastError();
        return false;
    }
    (void)device;
    g_write_value = (write_value_fn)w;
    g_wait_value = (wait_value_fn)q;
    g_memops_state = 1;
    return true;
}

int run_host(dsb_erk_plan *h, const dsb_erk_host_args *a, cudaStream_t stream)
{
    const dsb_erk_run_args &r = a->run;
    const int64_t n = r.n;
    if (!h->fused_launch_flow) {
        set_error("dsb_erk_run_host: no flow-controlled fused kernel for rhs_id=%d dim=%d stages=%d flags=%u (fast "
                  "flavour, <= 7 stages); use chunked dsb_erk_run launches", h->rhs_id, h->dim, h->stages, h->flags);
        return DSB_ERR_UNSUPPORTED;
    }
    // host_y0 == NULL: the initial state is already resident (run.y_init, or run.y in place) -- nothing is uploaded and only
    // the result side of the pipeline runs.  The destinations may then be ANY unified-address pointers, e.g. another
    // GPU's buffer mapped through dsb_peer_open: the gather of a sharded ensemble as chunked peer copies that overlap
    // the integration (desolver_b200/distributed.py, PushTarget).
    const bool upload = a->host_y0 != nullptr;
    if (n < 0 || n > 0x7fffffffLL || !r.y || !r.t || !r.dt || !r.counters || !r.accepted || !r.rejected || !r.status ||
        !a->host_y || !a->flags ||
        (upload && !a->stream_up) || !a->stream_down || a->chunk_shift < 5 || a->chunk_shift > 30 || !r.fresh ||
        (upload && r.y_init) || r.store.pool) {
        set_error("dsb_erk_run_host: bad argument (device buffers with fresh = 1, no dense output; with host_y0 no y_init; "
                  "result buffers, flags and the extra stream(s) are required; 5 <= chunk_shift <= 30)");
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (n == 0) return DSB_OK;
    if (!resolve_memops(h->device)) {
        set_error("dsb_erk_run_host: the driver does not provide cuStreamWriteValue32 / cuStreamWaitValue32");
        return DSB_ERR_UNSUPPORTED;
    }
    cudaStream_t up = (cudaStream_t)a->stream_up, down = (cudaStream_t)a->stream_down;
    const int64_t chunk = (int64_t)1 << a->chunk_shift;
    const int64_t nchunks = (n + chunk - 1) >> a->chunk_shift;
    const int64_t ld = r.stride > 0 ? r.stride : n, hld = a->host_stride > 0 ? a->host_stride : n;
    const int dim = h->dim;
    if (!h->ev_start) {
        DSB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
        DSB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    }
    DSB_CUDA_CHECK(cudaMemsetAsync(a->flags, 0, sizeof(uint32_t) * (size_t)(1 + nchunks), stream));
    DSB_CUDA_CHECK(cudaEventRecord(h->ev_start, stream));
    if (upload) DSB_CUDA_CHECK(cudaStreamWaitEvent(up, h->ev_start, 0));
    DSB_CUDA_CHECK(cudaStreamWaitEvent(down, h->ev_start, 0));

    // uploads: chunk c of every state row in one 2-D copy, then publish the resident count.  The kernel goes in
    // right after the first chunk has been queued, so that its start is not delayed by the remaining enqueues.
    int status = DSB_OK;
    if (!upload) {
        dsb_erk_run_args run = r;
        run.avail = nullptr;                       // everything is resident: the kernel never waits
        run.done = a->flags + 1;
        run.done_shift = a->chunk_shift;
        status = h->fused_launch_flow(h, &run, stream);
        if (status != DSB_OK) return status;
    }
    for (int64_t c = 0; upload && c < nchunks; ++c) {
        const int64_t lo = c << a->chunk_shift, len = (lo + chunk <= n ? chunk : n - lo);
        DSB_CUDA_CHECK(cudaMemcpy2DAsync(r.y + lo, sizeof(double) * ld, a->host_y0 + lo, sizeof(double) * hld,
                                         sizeof(double) * len, dim, cudaMemcpyHostToDevice, up));
        if (g_write_value((CUstream)up, (CUdeviceptr)(uintptr_t)a->flags, (cuuint32_t)(lo + len), 0) != CUDA_SUCCESS) {
            // a kernel that is already running gives up after its bounded wait and reports through counters[6]
            set_error("dsb_erk_run_host: cuStreamWriteValue32 failed");
            return DSB_ERR_CUDA;
        }
        if (c == 0) {
            dsb_erk_run_args run = r;
            run.avail = a->flags;
            run.done = a->flags + 1;
            run.done_shift = a->chunk_shift;
            status = h->fused_launch_flow(h, &run, stream);
            if (status != DSB_OK) return status;
        }
    }
    // downloads, each gated on its chunk's completion counter
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t lo = c << a->chunk_shift, len = (lo + chunk <= n ? chunk : n - lo);
        if (g_wait_value((CUstream)down, (CUdeviceptr)(uintptr_t)(a->flags + 1 + c), (cuuint32_t)len,
                         CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
            set_error("dsb_erk_run_host: cuStreamWaitValue32 failed");
            return DSB_ERR_CUDA;
        }
        DSB_CUDA_CHECK(cudaMemcpy2DAsync(a->host_y + lo, sizeof(double) * hld, r.y + lo, sizeof(double) * ld,
                                         sizeof(double) * len, dim, cudaMemcpyDefault, down));
        if (a->host_accepted)    // NULL (any of the three counter arrays): not downloaded, the caller reads run.* on demand
            DSB_CUDA_CHECK(cudaMemcpyAsync(a->host_accepted + lo, r.accepted + lo, sizeof(int32_t) * len, cudaMemcpyDefault, down));
        if (a->host_rejected)
            DSB_CUDA_CHECK(cudaMemcpyAsync(a->host_rejected + lo, r.rejected + lo, sizeof(int32_t) * len, cudaMemcpyDefault, down));
        if (a->host_status)      // NULL: the caller fetches the status words only if counters[4..5] report a failure
            DSB_CUDA_CHECK(cudaMemcpyAsync(a->host_status + lo, r.status + lo, sizeof(int32_t) * len, cudaMemcpyDefault, down));
        if (a->host_t)
            DSB_CUDA_CHECK(cudaMemcpyAsync(a->host_t + lo, r.t + lo, sizeof(double) * len, cudaMemcpyDefault, down));
    }
    DSB_CUDA_CHECK(cudaEventRecord(h->ev_done, down));
    DSB_CUDA_CHECK(cudaStreamWaitEvent(stream, h->ev_done, 0));
    return DSB_OK;
}

}  // namespace dsb
