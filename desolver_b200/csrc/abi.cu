// abi.cu -- the extern "C" surface of libdesolver_b200.so (include/desolver_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "plan.h"
#include <map>
#include <mutex>

namespace dsb {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int launch_dgemm(const double *A, const double *B, double *C, int64_t M, int64_t N, int64_t K, int sm_count, cudaStream_t stream);
int run_host(dsb_erk_plan *h, const dsb_erk_host_args *a, cudaStream_t stream);
int stage_begin(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int first_call, cudaStream_t stream);
int stage_state(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int stage, cudaStream_t stream);
int stage_finish(const dsb_erk_plan *h, const dsb_erk_stage_args *a, cudaStream_t stream);
int stage_finish_events(const dsb_erk_plan *h, const dsb_erk_stage_args *a, int phase, const double *f1, const double *roots,
                        cudaStream_t stream);

}  // namespace dsb

#define DSB_STR2(x) #x
#define DSB_STR(x) DSB_STR2(x)

using namespace dsb;

extern "C" {

int dsb_abi_version(void) { return DSB_ABI_VERSION; }

const char *dsb_build_info(void)
{
    return "libdesolver_b200 abi=3 arch=sm_100a nvcc=" DSB_STR(__CUDACC_VER_MAJOR__) "." DSB_STR(__CUDACC_VER_MINOR__)
           " fused_stages={1,2,4,6,7,13,17} rhs={lorenz,van_der_pol,harmonic_oscillator,forced,duffing}";
}

const char *dsb_last_error(void) { return g_error; }

// ---- right-hand sides compiled OUTSIDE this library (desolver_b200.rhs_plugins.register_device_rhs) ----------------
// A plug-in shared object instantiates the same kernel templates (erk_fused.cuh, plan.h) for its own Rhs struct and
// hands over the function that binds them to a plan; dsb_erk_create finds it here by its id.
namespace {
struct UserRhs { int dim; bool (*select)(dsb_erk_plan *); };
std::map<int, UserRhs> g_user_rhs;
std::mutex g_user_rhs_mutex;
}

int dsb_rhs_register(int rhs_id, int dim, void *select_fn)
{
    if (rhs_id < DSB_RHS_USER_BASE || dim < 1 || !select_fn) {
        set_error("dsb_rhs_register: rhs_id must be >= %d, dim >= 1, select_fn non-null (rhs_id=%d dim=%d)",
                  (int)DSB_RHS_USER_BASE, rhs_id, dim);
        return DSB_ERR_BAD_ARGUMENT;
    }
    std::lock_guard<std::mutex> lock(g_user_rhs_mutex);
    g_user_rhs[rhs_id] = UserRhs{dim, reinterpret_cast<bool (*)(dsb_erk_plan *)>(select_fn)};
    return DSB_OK;
}

int dsb_erk_create(dsb_erk_handle *out, int device, int rhs_id, int dim, int stages,
                   const double *tableau_intermediate, const double *tableau_final, int final_rows,
                   double order, unsigned flags)
{
    if (!out || !tableau_intermediate || !tableau_final || stages < 1 || (final_rows != 1 && final_rows != 2) ||
        dim < 1 || !(order > 0.0)) {
        set_error("dsb_erk_create: bad argument (stages=%d final_rows=%d dim=%d order=%g)", stages, final_rows, dim, order);
        return DSB_ERR_BAD_ARGUMENT;
    }
    // explicit methods only: a[s][j] must vanish for j >= s (integrator_types.py:135)
    for (int s = 0; s < stages; ++s)
        for (int j = s; j < stages; ++j)
            if (tableau_intermediate[s * (stages + 1) + 1 + j] != 0.0) {
                set_error("dsb_erk_create: tableau is implicit (a[%d][%d] != 0); only explicit RK is supported", s, j);
                return DSB_ERR_UNSUPPORTED;
            }
    DeviceGuard guard(device);                 // the caller's current device is restored on return
    int sm_count = 0;
    DSB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    dsb_erk_plan *p = new (std::nothrow) dsb_erk_plan;
    if (!p) { set_error("out of host memory"); return DSB_ERR_BAD_ARGUMENT; }
    p->device = device; p->rhs_id = rhs_id; p->dim = dim; p->stages = stages; p->final_rows = final_rows;
    p->order = order; p->flags = flags;
    p->inter.assign(tableau_intermediate, tableau_intermediate + (size_t)stages * (stages + 1));
    p->fin.assign(tableau_final, tableau_final + (size_t)final_rows * (stages + 1));
    p->sm_count = sm_count;

    bool fused = false;
    switch (rhs_id) {
    case DSB_RHS_LORENZ: fused = dim == 3 && select_fused_lorenz(p); break;
    case DSB_RHS_VDP:    fused = dim == 2 && select_fused_vdp(p); break;
    case DSB_RHS_SHO:    fused = dim == 2 && select_fused_sho(p); break;
    case DSB_RHS_FORCED: fused = dim == 2 && select_fused_forced(p); break;
    case DSB_RHS_DUFFING: fused = dim == 2 && select_fused_duffing(p); break;
    default: {
        std::lock_guard<std::mutex> lock(g_user_rhs_mutex);
        auto it = g_user_rhs.find(rhs_id);
        if (it != g_user_rhs.end()) fused = dim == it->second.dim && it->second.select(p);
        break;
    }
    }
    if (!fused) p->fused_launch = p->fused_launch_flow = p->fused_launch_events = p->fused_launch_events_dy =
        p->fused_launch_scatter = p->fused_launch_dense = p->fused_launch_multi = nullptr;

    double2 host_tab[fm::ATAN_TABLE_SIZE];
    fm::build_atan_table(host_tab);
    if (cudaMalloc(&p->d_atan, sizeof(host_tab)) != cudaSuccess ||
        cudaMemcpy(p->d_atan, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("dsb_erk_create: device table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete p;
        return DSB_ERR_CUDA;
    }
    if (final_rows == 2) {
        // step-size factor of a step's first attempt, tabulated for this method's order (factor_table.h)
        std::vector<ft::Entry> host_ft(ft::ENTRIES);
        ft::build_table(order, 0.8, host_ft.data());
        if (cudaMalloc(&p->d_factor, sizeof(ft::Entry) * ft::ENTRIES) != cudaSuccess ||
            cudaMemcpy(p->d_factor, host_ft.data(), sizeof(ft::Entry) * ft::ENTRIES, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("dsb_erk_create: device table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            cudaFree(p->d_atan);
            delete p;
            return DSB_ERR_CUDA;
        }
    }
    *out = p;
    return DSB_OK;
}

int dsb_erk_destroy(dsb_erk_handle h)
{
    if (!h) return DSB_OK;
    DeviceGuard guard(h->device);
    if (h->d_atan) cudaFree(h->d_atan);
    if (h->d_tab) cudaFree(h->d_tab);
    if (h->d_factor) cudaFree(h->d_factor);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    delete h;
    return DSB_OK;
}

int dsb_erk_run(dsb_erk_handle h, const dsb_erk_run_args *a, void *stream)
{
    if (!h || !a) { set_error("dsb_erk_run: null handle or args"); return DSB_ERR_BAD_ARGUMENT; }
    if (!h->fused_launch) {
        set_error("dsb_erk_run: no fused kernel for rhs_id=%d dim=%d stages=%d; use the stage-level entry points",
                  h->rhs_id, h->dim, h->stages);
        return DSB_ERR_UNSUPPORTED;
    }
    if (a->n < 0 || !a->y || !a->t || !a->dt || !a->counters) {
        set_error("dsb_erk_run: y, t, dt and counters are required");
        return DSB_ERR_BAD_ARGUMENT;
    }
    const dsb_node_store &ns = a->store;
    if (ns.pool && (!ns.cursor || !ns.page_fill || !ns.node_count || ns.width != 32 || ns.page_entries < 32 ||
                    ns.page_entries % 32 != 0 || ns.capacity < ns.page_entries || ns.capacity % ns.page_entries != 0 ||
                    ns.dim != h->dim || a->out_y || a->n_events > 0)) {
        set_error("dsb_erk_run: the node store of a fused run needs cursor, page_fill, node_count, width = 32, page_entries a "
                  "multiple of 32, capacity a multiple of page_entries, dim = %d, and neither events nor a scatter-store in "
                  "the same call", h->dim);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (a->stride != 0 && a->stride < a->n) {
        set_error("dsb_erk_run: stride (%lld) must be 0 or >= n (%lld)", (long long)a->stride, (long long)a->n);
        return DSB_ERR_BAD_ARGUMENT;
    }
    if (a->n == 0) return DSB_OK;
    DeviceGuard guard(h->device);
    if (a->n_events > 0) {
        if (a->n_events > DSB_MAX_EVENTS || !a->ev_coef || !a->ev_flags || !a->ev_records || !a->ev_count ||
            a->ev_capacity < 1 || a->store.pool || a->stops) {
            set_error("dsb_erk_run: events need ev_coef, ev_flags, ev_records, ev_count, ev_capacity >= 1, at most %d "
                      "events, and no dense-output store in the same call", DSB_MAX_EVENTS);
            return DSB_ERR_BAD_ARGUMENT;
        }
        if (!h->fused_launch_events) {
            set_error("dsb_erk_run: no event-detecting kernel for rhs_id=%d stages=%d (<= 7 stages)", h->rhs_id, h->stages);
            return DSB_ERR_UNSUPPORTED;
        }
        return (a->ev_coef_dy ? h->fused_launch_events_dy : h->fused_launch_events)(h, a, (cudaStream_t)stream);
    }
    if (a->out_y) {
        if (!a->out_t || !a->out_dt || !a->out_accepted || !a->out_rejected || !a->out_status || !a->fresh || !a->y_init ||
            a->out_index_mul < 1 || a->out_index_add < 0 || a->out_stride < 1 || a->store.pool || a->stops) {
            set_error("dsb_erk_run: the scatter-store needs every out_* pointer, out_stride >= 1, out_index_mul >= 1, "
                      "out_index_add >= 0, fresh != 0, y_init, and no dense-output store");
            return DSB_ERR_BAD_ARGUMENT;
        }
        if (!h->fused_launch_scatter) {
            set_error("dsb_erk_run: no scatter-store kernel for rhs_id=%d stages=%d flags=%u (fast flavour, <= 7 stages)",
                      h->rhs_id, h->stages, h->flags);
            return DSB_ERR_UNSUPPORTED;
        }
        return h->fused_launch_scatter(h, a, (cudaStream_t)stream);
    }
    if (a->stops) {
        if (a->n_stops < 1 || !a->stop_y || a->stop_stride < a->n || a->store.pool || a->max_steps >= 0) {
            set_error("dsb_erk_run: a stop list needs n_stops >= 1, stop_y, stop_stride >= n, max_steps < 0 and neither events, "
                      "a node store nor a scatter-store in the same call");
            return DSB_ERR_BAD_ARGUMENT;
        }
        if (!h->fused_launch_multi) {
            set_error("dsb_erk_run: no multi-stop kernel for rhs_id=%d stages=%d (<= 7 stages)", h->rhs_id, h->stages);
            return DSB_ERR_UNSUPPORTED;
        }
        return h->fused_launch_multi(h, a, (cudaStream_t)stream);
    }
    if (a->store.pool) {
        if (!h->fused_launch_dense) { set_error("dsb_erk_run: no dense-output kernel for this handle"); return DSB_ERR_UNSUPPORTED; }
        return h->fused_launch_dense(h, a, (cudaStream_t)stream);
    }
    return h->fused_launch(h, a, (cudaStream_t)stream);
}

int dsb_erk_run_host(dsb_erk_handle h, const dsb_erk_host_args *a, void *stream)
{
    if (!h || !a) { set_error("dsb_erk_run_host: null handle or args"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return run_host(h, a, (cudaStream_t)stream);
}

int dsb_erk_run_launches(dsb_erk_handle h) { return (h && h->fused_launch) ? 1 : 0; }

int dsb_erk_stage_begin(dsb_erk_handle h, const dsb_erk_stage_args *a, int first_call, void *stream)
{
    if (!h) { set_error("dsb_erk_stage_begin: null handle"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return stage_begin(h, a, first_call, (cudaStream_t)stream);
}

int dsb_erk_stage_state(dsb_erk_handle h, const dsb_erk_stage_args *a, int stage, void *stream)
{
    if (!h) { set_error("dsb_erk_stage_state: null handle"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return stage_state(h, a, stage, (cudaStream_t)stream);
}

int dsb_erk_stage_finish(dsb_erk_handle h, const dsb_erk_stage_args *a, void *stream)
{
    if (!h) { set_error("dsb_erk_stage_finish: null handle"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return stage_finish(h, a, (cudaStream_t)stream);
}

int dsb_erk_stage_finish_events(dsb_erk_handle h, const dsb_erk_stage_args *a, int phase, const double *f1, void *stream)
{
    if (!h) { set_error("dsb_erk_stage_finish_events: null handle"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return stage_finish_events(h, a, phase, f1, nullptr, (cudaStream_t)stream);
}

int dsb_erk_stage_finish_events_located(dsb_erk_handle h, const dsb_erk_stage_args *a, const double *f1, const double *roots,
                                        void *stream)
{
    if (!h || !roots) { set_error("dsb_erk_stage_finish_events_located: null handle or roots"); return DSB_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    return stage_finish_events(h, a, 1, f1, roots, (cudaStream_t)stream);
}

int dsb_erk_stage_launches(dsb_erk_handle h) { return h ? 1 + h->stages + 3 : 0; }

int dsb_dgemm(const double *A, const double *B, double *C, int64_t M, int64_t N, int64_t K, void *stream)
{
    if (!A || !B || !C) { set_error("dsb_dgemm: null pointer"); return DSB_ERR_BAD_ARGUMENT; }
    int dev = 0, sms = 0;
    DSB_CUDA_CHECK(cudaGetDevice(&dev));
    DSB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return launch_dgemm(A, B, C, M, N, K, sms, (cudaStream_t)stream);
}

}  // extern "C"
