// event_search.cu -- bracketed root search for ARBITRARY event functions.
//
// The reference locates an event inside an accepted step by running a vectorised Brent iteration on
// t -> event(t, sol(t)[, sol.grad(t)], **constants) over [t_prev, t_next] (differential_system.py:60-167,
// utilities/optimizer.py:178-366).  An affine event is solved in closed form inside the step kernels (the cubic of
// erk_fused.cuh / erk_stage.cu); for any other callable the function values have to come from the caller's code, so the
// search is split into kernels with the callable evaluated between them, batched over every (event, trajectory) pair:
//
//   dsb_event_search_begin   brackets [0, 1] in step fractions where g changes sign over an accepted step in an allowed
//                            direction (fa * fb < 0, optimizer.py:252), first false-position iterate
//   loop:  tq = t0 + x (t1 - t0)  ->  y(tq) by dsb_hermite_eval  ->  fx = event(tq, y(tq), ...)   [caller, same stream]
//          dsb_event_search_step     Illinois update of the bracket (regula falsi with the stale end's value halved;
//                                    superlinear, never leaves the bracket), next iterate, convergence flag
//   dsb_event_search_roots   root times (NaN where no event), handed to the event kernels' record / order / terminal logic
//                            (dsb_erk_stage_finish_events_located, dsb_symp_stage_commit_events_located)
#include "common.cuh"
#include <math_constants.h>

namespace dsb {

__device__ __forceinline__ double false_position(double a, double b, double fa, double fb)
{
    double c = (fa * b - fb * a) / (fa - fb);
    if (!(c > fmin(a, b) && c < fmax(a, b))) c = 0.5 * (a + b);         // also catches inf / nan
    return c;
}

// state per (event e, trajectory i), all [n_events][n]: lo, hi (step fractions), flo, fhi, x (iterate), side (int)
__global__ void event_search_begin_kernel(dsb_event_search s, const double *__restrict__ g0, const double *__restrict__ g1,
                                          const int32_t *__restrict__ flags, int need_mask, int skip_mask,
                                          const int32_t *__restrict__ ev_flags)
{
    const long long total = (long long)s.n_events * s.n;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long i = q % s.n;
        const int e = (int)(q / s.n);
        const int fl = flags[i];
        const double a = g0[q], b = g1[q];
        const int dir = ev_flags[e] & 3;
        const bool stepped = (fl & need_mask) == need_mask && !(fl & skip_mask);
        const bool rising = a < 0.0;
        const bool hit = stepped && a * b < 0.0 && (dir == 0 || (dir == 1) == rising);
        s.lo[q] = 0.0; s.hi[q] = 1.0; s.flo[q] = a; s.fhi[q] = b;
        s.side[q] = hit ? 0 : -100;                                       // -100: no event in this step; 100: converged
        const double x = hit ? false_position(0.0, 1.0, a, b) : 0.0;
        s.x[q] = x;
        s.tq[q] = fma(x, s.t1[i] - s.t0[i], s.t0[i]);
        if (hit) atomicAdd((unsigned long long *)s.counter, 1ull);        // searches in flight (host: skip the loop if 0)
    }
}

__global__ void event_search_step_kernel(dsb_event_search s, const double *__restrict__ fx_all)
{
    const long long total = (long long)s.n_events * s.n;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        int side = s.side[q];
        if (side == -100 || side == 100) continue;
        const long long i = q % s.n;
        double a = s.lo[q], b = s.hi[q], fa = s.flo[q], fb = s.fhi[q];
        const double c = s.x[q], fc = fx_all[q];
        if (fc == 0.0 || !(fc == fc)) {                                   // exact root (or the callable went non-finite: stop here)
            s.side[q] = 100;
            atomicAdd((unsigned long long *)s.counter, ~0ull);
            continue;
        }
        if ((fc < 0.0) == (fb < 0.0)) {                                   // c replaces b
            b = c; fb = fc;
            if (side == -1) fa *= 0.5;
            side = -1;
        } else {                                                          // c replaces a
            a = c; fa = fc;
            if (side == 1) fb *= 0.5;
            side = 1;
        }
        double x = false_position(a, b, fa, fb);
        const bool done = fabs(b - a) <= 4.5e-16 || x == c;
        if (done) {
            // the end whose (unscaled) residual is smaller; both are within one spacing of the root
            x = fabs(fa) < fabs(fb) ? a : b;
            side = 100;
            atomicAdd((unsigned long long *)s.counter, ~0ull);
        }
        s.lo[q] = a; s.hi[q] = b; s.flo[q] = fa; s.fhi[q] = fb; s.side[q] = side; s.x[q] = x;
        s.tq[q] = fma(x, s.t1[i] - s.t0[i], s.t0[i]);
    }
}

__global__ void event_search_roots_kernel(dsb_event_search s, double *__restrict__ roots)
{
    const long long total = (long long)s.n_events * s.n;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
        roots[q] = s.side[q] == -100 ? CUDART_NAN : s.tq[q];
}

static int search_grid(const dsb_event_search *s)
{
    long long b = ((long long)s->n_events * s->n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

static int check_search(const dsb_event_search *s, const char *who)
{
    if (!s || s->n < 1 || s->n_events < 1 || !s->lo || !s->hi || !s->flo || !s->fhi || !s->x || !s->tq || !s->side || !s->t0 ||
        !s->t1 || !s->counter) {
        set_error("%s: bad argument (every array of dsb_event_search is required)", who);
        return DSB_ERR_BAD_ARGUMENT;
    }
    return DSB_OK;
}

}  // namespace dsb

using namespace dsb;

extern "C" {

int dsb_event_search_begin(const dsb_event_search *s, const double *g0, const double *g1, const int32_t *flags, int32_t need_mask,
                           int32_t skip_mask, const int32_t *ev_flags, void *stream)
{
    int rc = check_search(s, "dsb_event_search_begin");
    if (rc != DSB_OK) return rc;
    if (!g0 || !g1 || !flags || !ev_flags) { set_error("dsb_event_search_begin: null argument"); return DSB_ERR_BAD_ARGUMENT; }
    event_search_begin_kernel<<<search_grid(s), 256, 0, (cudaStream_t)stream>>>(*s, g0, g1, flags, need_mask, skip_mask, ev_flags);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_event_search_step(const dsb_event_search *s, const double *fx, void *stream)
{
    int rc = check_search(s, "dsb_event_search_step");
    if (rc != DSB_OK) return rc;
    if (!fx) { set_error("dsb_event_search_step: null fx"); return DSB_ERR_BAD_ARGUMENT; }
    event_search_step_kernel<<<search_grid(s), 256, 0, (cudaStream_t)stream>>>(*s, fx);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

int dsb_event_search_roots(const dsb_event_search *s, double *roots, void *stream)
{
    int rc = check_search(s, "dsb_event_search_roots");
    if (rc != DSB_OK) return rc;
    if (!roots) { set_error("dsb_event_search_roots: null roots"); return DSB_ERR_BAD_ARGUMENT; }
    event_search_roots_kernel<<<search_grid(s), 256, 0, (cudaStream_t)stream>>>(*s, roots);
    DSB_CUDA_CHECK(cudaGetLastError());
    return DSB_OK;
}

}  // extern "C"
