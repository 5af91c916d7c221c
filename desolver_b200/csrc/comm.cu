// comm.cu -- the collectives of the sharded ensemble behind the C ABI (include/desolver_b200.h, "collectives"):
// the termination / statistics all-reduce and the final gather of round-robin shards, over NCCL (NVLink 5 /
// NVSwitch), plus CUDA-IPC peer buffers for the scatter-store variant of the fused kernel (results written straight
// into the destination rank's memory over NVLink while the integration is still running).
//
// NCCL is bound at RUN time (dlopen of the libnccl.so.2 that the process already carries -- torch's -- or the system
// one): a single-GPU user never needs it, and the library keeps linking against cudart only.
#include <dlfcn.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace dsb {

// ---- the slice of the NCCL API that is used (stable across NCCL 2.x) ----------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // 0 = ncclSuccess
enum { NCCL_INT8 = 0, NCCL_UINT64 = 5 };  // ncclDataType_t
enum { NCCL_SUM = 0 };                    // ncclRedOp_t

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static int g_nccl_state = 0;              // 0 unknown, 1 bound, -1 unavailable

static bool bind_nccl()
{
    if (g_nccl_state != 0) return g_nccl_state > 0;
    g_nccl_state = -1;
    // the copy the process already holds (torch links its own libnccl.so.2) wins over a second one
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { set_error("NCCL is not available: %s", dlerror()); return false; }
    NcclApi a;
    a.lib = lib;
#define DSB_SYM(field, name)                                                             \
    *(void **)(&a.field) = dlsym(lib, name);                                             \
    if (!a.field) { set_error("NCCL symbol %s not found", name); return false; }
    DSB_SYM(GetUniqueId, "ncclGetUniqueId")
    DSB_SYM(CommInitRank, "ncclCommInitRank")
    DSB_SYM(CommDestroy, "ncclCommDestroy")
    DSB_SYM(CommCount, "ncclCommCount")
    DSB_SYM(CommUserRank, "ncclCommUserRank")
    DSB_SYM(AllReduce, "ncclAllReduce")
    DSB_SYM(AllGather, "ncclAllGather")
    DSB_SYM(GetErrorString, "ncclGetErrorString")
    DSB_SYM(GetVersion, "ncclGetVersion")
#undef DSB_SYM
    g_nccl = a;
    g_nccl_state = 1;
    return true;
}

#define DSB_NCCL_CHECK(expr)                                                              \
    do {                                                                                  \
        ncclResult_t r__ = (expr);                                                        \
        if (r__ != 0) {                                                                   \
            dsb::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            return DSB_ERR_NCCL;                                                          \
        }                                                                                 \
    } while (0)

}  // namespace dsb

struct dsb_comm {
    dsb::ncclComm_t comm = nullptr;
    int device = 0, nranks = 1, rank = 0;
    bool owned = false;                   // created by dsb_comm_init_rank (destroyed with the handle)
};

namespace dsb {

// Interleave of the gathered shards: full[r][j * W + w] = staged[w][r][j]  (trajectory i lives on rank i mod W).
// grid.y = row; a thread owns one destination element: the writes of a warp are one contiguous segment, its reads are
// W streams of 32 / W consecutive elements each (whole sectors for W <= 8 doubles).  WSHIFT >= 0: W = 2^WSHIFT, the
// index split is a shift and a mask (an integer division per element costs more than the memory traffic).
template <class T, int WSHIFT>
__global__ void interleave_kernel(const T *__restrict__ staged, T *__restrict__ full, int rows, long long n_max,
                                  long long n_global, int W)
{
    const int r = blockIdx.y;
    const T *src = staged + (long long)r * n_max;
    T *dst = full + (long long)r * n_global;
    const long long slot = (long long)rows * n_max;              // elements per rank block
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_global; i += (long long)gridDim.x * blockDim.x) {
        long long j;
        int w;
        if constexpr (WSHIFT >= 0) { j = i >> WSHIFT; w = (int)(i & ((1 << WSHIFT) - 1)); }
        else { j = i / W; w = (int)(i - j * W); }
        dst[i] = src[w * slot + j];
    }
}

template <class T>
static void launch_interleave(const T *staged, T *full, int rows, long long n_max, long long n_global, int W, int sms, cudaStream_t st)
{
    long long bx = (n_global + 255) / 256;
    const long long cap = ((long long)sms * 16 + rows - 1) / rows;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)bx, (unsigned)rows);
    int shift = -1;
    for (int k = 0; k < 16; ++k) if ((1 << k) == W) shift = k;
    switch (shift) {
    case 0: interleave_kernel<T, 0><<<grid, 256, 0, st>>>(staged, full, rows, n_max, n_global, W); break;
    case 1: interleave_kernel<T, 1><<<grid, 256, 0, st>>>(staged, full, rows, n_max, n_global, W); break;
    case 2: interleave_kernel<T, 2><<<grid, 256, 0, st>>>(staged, full, rows, n_max, n_global, W); break;
    case 3: interleave_kernel<T, 3><<<grid, 256, 0, st>>>(staged, full, rows, n_max, n_global, W); break;
    default: interleave_kernel<T, -1><<<grid, 256, 0, st>>>(staged, full, rows, n_max, n_global, W); break;
    }
}

// Pads a shard [rows][n_local] to [rows][n_max] inside this rank's slot of the staging buffer (zeros beyond n_local).
template <class T>
__global__ void pad_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, int rows, long long n_local, long long n_max)
{
    const long long total = (long long)rows * n_max;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / n_max, j = e - r * n_max;
        dst[e] = j < n_local ? src[r * n_local + j] : T(0);
    }
}

static int grid_1d(long long total, int threads, int sms)
{
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)sms * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

template <class T>
static int gather_typed(dsb_comm *c, const void *shard, void *staging, void *full, int rows, long long n_local,
                        long long n_global, int sms, cudaStream_t st)
{
    const int W = c->nranks;
    const long long n_max = (n_global + W - 1) / W;
    const size_t slot = (size_t)rows * n_max * sizeof(T);
    T *mine = (T *)((char *)staging + slot * c->rank);
    if (n_local == n_max) {
        DSB_CUDA_CHECK(cudaMemcpyAsync(mine, shard, slot, cudaMemcpyDeviceToDevice, st));
    } else {
        pad_rows_kernel<T><<<grid_1d((long long)rows * n_max, 256, sms), 256, 0, st>>>((const T *)shard, mine, rows, n_local, n_max);
        DSB_CUDA_CHECK(cudaGetLastError());
    }
    if (W > 1) DSB_NCCL_CHECK(g_nccl.AllGather(mine, staging, slot, NCCL_INT8, c->comm, st));   // in place
    launch_interleave<T>((const T *)staging, (T *)full, rows, n_max, n_global, W, sms, st);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // namespace dsb

using namespace dsb;

extern "C" {

int dsb_comm_available(void) { return bind_nccl() ? 1 : 0; }

int dsb_comm_nccl_version(void)
{
    int v = 0;
    if (!bind_nccl() || g_nccl.GetVersion(&v) != 0) return 0;
    return v;
}

int dsb_comm_get_unique_id(void *id_out)
{
    if (!id_out) { set_error("dsb_comm_get_unique_id: null output"); return DSB_ERR_BAD_ARGUMENT; }
    if (!bind_nccl()) return DSB_ERR_UNSUPPORTED;
    ncclUniqueId id;
    DSB_NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return DSB_OK;
}

int dsb_comm_init_rank(dsb_comm_handle *out, int device, int nranks, int rank, const void *id)
{
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("dsb_comm_init_rank: bad argument (nranks=%d rank=%d)", nranks, rank);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (!bind_nccl()) return DSB_ERR_UNSUPPORTED;
    DeviceGuard guard(device);
    dsb_comm *c = new (std::nothrow) dsb_comm;
    if (!c) { set_error("out of host memory"); return DSB_ERR_BAD_ARGUMENT; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return DSB_ERR_NCCL;
    }
    c->device = device; c->nranks = nranks; c->rank = rank; c->owned = true;
    *out = c;
    return DSB_OK;
}

int dsb_comm_from_nccl(dsb_comm_handle *out, void *nccl_comm, int device)
{
    if (!out || !nccl_comm) { set_error("dsb_comm_from_nccl: null argument"); return DSB_ERR_BAD_ARGUMENT; }
    if (!bind_nccl()) return DSB_ERR_UNSUPPORTED;
    dsb_comm *c = new (std::nothrow) dsb_comm;
    if (!c) { set_error("out of host memory"); return DSB_ERR_BAD_ARGUMENT; }
    c->comm = (ncclComm_t)nccl_comm;
    c->device = device;
    c->owned = false;
    if (g_nccl.CommCount(c->comm, &c->nranks) != 0 || g_nccl.CommUserRank(c->comm, &c->rank) != 0) {
        set_error("dsb_comm_from_nccl: not a valid ncclComm_t");
        delete c;
        return DSB_ERR_NCCL;
    }
    *out = c;
    return DSB_OK;
}

int dsb_comm_destroy(dsb_comm_handle c)
{
    if (!c) return DSB_OK;
    if (c->owned && c->comm && g_nccl_state > 0) {
        DeviceGuard guard(c->device);
        g_nccl.CommDestroy(c->comm);
    }
    delete c;
    return DSB_OK;
}

int dsb_comm_size(dsb_comm_handle c) { return c ? c->nranks : 0; }
int dsb_comm_rank(dsb_comm_handle c) { return c ? c->rank : -1; }

int dsb_comm_allreduce_counters(dsb_comm_handle c, const uint64_t *send, uint64_t *recv, int count, void *stream)
{
    if (!c || !send || !recv || count < 1) { set_error("dsb_comm_allreduce_counters: bad argument"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(c->device);
    DSB_NCCL_CHECK(g_nccl.AllReduce(send, recv, (size_t)count, NCCL_UINT64, NCCL_SUM, c->comm, (cudaStream_t)stream));
    return DSB_OK;
}

int dsb_comm_gather(dsb_comm_handle c, const void *shard, void *staging, void *full, int rows, int elem_bytes,
                    int64_t n_local, int64_t n_global, void *stream)
{
    if (!c || !shard || !staging || !full || rows < 1 || (elem_bytes != 4 && elem_bytes != 8) || n_local < 0 || n_global < 1) {
        set_error("dsb_comm_gather: bad argument (rows=%d elem_bytes=%d n_local=%lld n_global=%lld)", rows, elem_bytes,
                  (long long)n_local, (long long)n_global);
        return DSB_ERR_BAD_ARGUMENT;
    }
    const long long expect = (n_global - c->rank + c->nranks - 1) / c->nranks;
    if (n_local != expect) {
        set_error("dsb_comm_gather: rank %d of %d holds %lld trajectories of %lld, expected %lld (round-robin)", c->rank,
                  c->nranks, (long long)n_local, (long long)n_global, expect);
        return DSB_ERR_BAD_ARGUMENT;
    }
    DeviceGuard guard(c->device);
    int sms = 0;
    DSB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (elem_bytes == 8)
        return gather_typed<unsigned long long>(c, shard, staging, full, rows, n_local, n_global, sms, (cudaStream_t)stream);
    return gather_typed<unsigned int>(c, shard, staging, full, rows, n_local, n_global, sms, (cudaStream_t)stream);
}

int dsb_interleave_shards(const void *staged, void *full, int rows, int elem_bytes, int64_t n_global, int world, void *stream)
{
    if (!staged || !full || rows < 1 || (elem_bytes != 4 && elem_bytes != 8) || n_global < 1 || world < 1) {
        set_error("dsb_interleave_shards: bad argument");
        return DSB_ERR_BAD_ARGUMENT;
    }
    int dev = 0, sms = 0;
    DSB_CUDA_CHECK(cudaGetDevice(&dev));
    DSB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long n_max = (n_global + world - 1) / world;
    if (elem_bytes == 8)
        launch_interleave<unsigned long long>((const unsigned long long *)staged, (unsigned long long *)full, rows, n_max, n_global,
                                              world, sms, (cudaStream_t)stream);
    else
        launch_interleave<unsigned int>((const unsigned int *)staged, (unsigned int *)full, rows, n_max, n_global, world, sms,
                                        (cudaStream_t)stream);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

// ---- CUDA-IPC peer buffers -----------------------------------------------------------------------------------
int dsb_peer_alloc(void **dev_ptr, int64_t bytes, void *ipc_handle_out, int device)
{
    if (!dev_ptr || !ipc_handle_out || bytes < 1) { set_error("dsb_peer_alloc: bad argument"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(device);
    void *p = nullptr;
    DSB_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaFree(p);
        return DSB_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == DSB_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return DSB_OK;
}

int dsb_peer_free(void *dev_ptr, int device)
{
    if (!dev_ptr) return DSB_OK;
    DeviceGuard guard(device);
    DSB_CUDA_CHECK(cudaFree(dev_ptr));
    return DSB_OK;
}

int dsb_peer_open(void **dev_ptr, const void *ipc_handle, int device)
{
    if (!dev_ptr || !ipc_handle) { set_error("dsb_peer_open: bad argument"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    DSB_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DSB_OK;
}

int dsb_peer_close(void *dev_ptr, int device)
{
    if (!dev_ptr) return DSB_OK;
    DeviceGuard guard(device);
    DSB_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return DSB_OK;
}

}  // extern "C"
