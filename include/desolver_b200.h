/*
 * desolver_b200.h -- C ABI of libdesolver_b200.so: the B200 (sm_100a) ensemble-integration
 * engine that sits behind desolver's Python API.
 *
 * The reference (Microno95/desolver v5.1.0, pure Python) has no FFI; its seams are Python
 * protocols.  Each entry point below states which reference code it
 * replaces (paths relative to /root/reference/desolver/).  Conventions:
 *   - plain C, no C++ types or exceptions cross the boundary;
 *   - every call returns a dsb_status (0 = OK); dsb_last_error() gives a thread-local message;
 *   - all array arguments are BORROWED raw DEVICE pointers owned by the caller (torch tensors
 *     on the Python side); the library never allocates, frees or resizes them;
 *   - every launch goes to the explicit `stream` argument (a cudaStream_t passed as void*);
 *     no call synchronises the device or the stream;
 *   - ensemble state is SoA: y[dim][n] (trajectory index fastest), t[n], dt[n];
 *   - tableaux are passed in from the Python class data (cls.tableau_intermediate /
 *     tableau_final / __order__), so there is one source of truth for coefficients.
 * One handle per (device, method, right-hand side); handles are not thread-safe.
 */
#ifndef DESOLVER_B200_H
#define DESOLVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB_ABI_VERSION 3

#if defined(__GNUC__)
#define DSB_API __attribute__((visibility("default")))
#else
#define DSB_API
#endif

typedef enum {
    DSB_OK = 0,
    DSB_ERR_BAD_ARGUMENT = 1,
    DSB_ERR_CUDA = 2,
    DSB_ERR_UNSUPPORTED = 3,  /* no fused kernel for this (rhs, method, dim) combination */
    DSB_ERR_NCCL = 4          /* a collective failed (dsb_comm_*)                                */
} dsb_status;

/* Built-in fused right-hand sides (desolver_b200/rhs_plugins.py tags the Python callables). */
typedef enum {
    DSB_RHS_LINEAR = 0,  /* y' = A y: stage-level entry points, A @ Y_stage as one FP64 GEMM */
    DSB_RHS_LORENZ = 1,  /* params: sigma, rho, beta                                          */
    DSB_RHS_VDP    = 2,  /* params: mu                                                        */
    DSB_RHS_NBODY  = 3,  /* state (2, nb, 3) per system; shared: G, eps2, masses              */
    DSB_RHS_SHO    = 4,  /* params: k, m                                                      */
    DSB_RHS_FORCED = 5,  /* y'' + y = t (time dependent; the reference's test fixture), no params */
    DSB_RHS_DUFFING = 6, /* x'' + delta x' + alpha x + beta x^3 = gamma cos(omega t); params in that order */
    DSB_RHS_EXTERNAL = 100 /* arbitrary torch callable evaluated between stage kernels       */
} dsb_rhs_id;

/* Per-trajectory status word written by the kernels.  The Python side maps
 * DSB_TRAJ_TOL_FAIL to exception_types.FailedToMeetTolerances wrapped in FailedIntegration
 * (differential_system.py:1089-1093, integrators/integrator_types.py:220-224). */
typedef enum {
    DSB_TRAJ_DONE = 0,      /* reached tf                                                     */
    DSB_TRAJ_TOL_FAIL = 1,  /* 64 retries exhausted; state frozen at the failed step's start  */
    DSB_TRAJ_RUNNING = 2,   /* step budget of this call exhausted before tf (chunked runs)    */
    DSB_TRAJ_EVENT = 3      /* stopped at the root of a terminal event (reference: integration_status 2,
                               differential_system.py:1053-1055)                               */
} dsb_traj_status;

#define DSB_FLAG_STRICT 1u  /* reference operation order, no FMA contraction, libm pow/atan   */

#define DSB_MAX_PARAMS 6
#define DSB_MAX_EVENTS 4          /* fused kernels (dsb_erk_run)                                       */
#define DSB_MAX_STAGE_EVENTS 16   /* stage-level kernels (dsb_erk_stage_*, dsb_symp_stage_*)           */
#define DSB_MAX_FUSED_STAGES 17

/* ---- library ------------------------------------------------------------------------- */
DSB_API int         dsb_abi_version(void);
DSB_API const char *dsb_build_info(void);      /* contains "sm_100a"                                 */
DSB_API const char *dsb_last_error(void);

/* ---- fused explicit Runge-Kutta ensemble engine ----------------------------------------
 * Replaces, for a built-in right-hand side, the whole of OdeSystem.integrate's while loop
 * (differential_system.py:996-1084) together with RungeKuttaIntegrator.__call__/step/
 * get_error_estimate (integrators/integrator_types.py:162-228, 267-323), compute_step
 * (integrators/components/runge_kutta_methods.py:44-54) and update_timestep
 * (integrators/integrator_template.py:46-93), once per trajectory, in ONE persistent launch. */
typedef struct dsb_erk_plan *dsb_erk_handle;

DSB_API int dsb_erk_create(dsb_erk_handle *out, int device, int rhs_id, int dim, int stages,
                   const double *tableau_intermediate, /* stages x (stages+1), host */
                   const double *tableau_final,        /* final_rows x (stages+1), host */
                   int final_rows, double order, unsigned flags);
DSB_API int dsb_erk_destroy(dsb_erk_handle h);

/* Right-hand sides compiled outside this library.  The reference takes ANY Python callable (differential_system.py:
 * 338-344); a callable the caller also provides as CUDA source is compiled, with the kernel templates of this library
 * (csrc/erk_fused.cuh, csrc/plan.h), into a plug-in shared object of its own and registered here under an id
 * >= DSB_RHS_USER_BASE; dsb_erk_create(rhs_id = that id) then binds the persistent fused kernels of THAT right-hand side
 * (desolver_b200.rhs_plugins.register_device_rhs builds the plug-in with nvcc and calls this).  `select_fn` is the
 * plug-in's `bool select(dsb_erk_plan *)`: it must come from a build against the same headers as this library. */
#define DSB_RHS_USER_BASE 1000
DSB_API int dsb_rhs_register(int rhs_id, int dim, void *select_fn);

/* Dense output / per-step history: the paged structure-of-arrays node store (csrc/node_store.cu).  A node is
 * (trajectory, ordinal k, t, y[dim], f[dim] = rhs(t, y)) -- what the reference keeps per accepted step as one
 * CubicHermiteInterp object (integrators/integrator_types.py:109-119, differential_system.py:1020-1023) and as one row
 * of `a.t` / `a.y` (differential_system.py:1015-1018).  The pool is a log of entries in groups of `width`; group g
 * holds 2 + 2*dim rows of `width` doubles: row 0 = header {int32 trajectory (-1: empty), int32 k}, row 1 = t, rows
 * 2..1+dim = y, rows 2+dim.. = f; entry e is slot e % width of group e / width.  Everything is caller-owned device
 * memory; cursor and page_fill must be zeroed when the store is created (or recycled). */
typedef struct {
    double *pool;              /* NULL = no node store                                          */
    int64_t capacity;          /* entries the pool holds, a multiple of width (and of page_entries) */
    int32_t width;             /* 32: written by the fused kernels (any trajectory in any slot);
                                  n: written by dsb_node_append_rows (slot = trajectory)        */
    int32_t page_entries;      /* fused kernels: a warp claims this many consecutive entries per atomic
                                  (multiple of 32, >= 32)                                       */
    int32_t dim;
    uint64_t *cursor;          /* [4]: [0] entries claimed so far, [1] nodes dropped because the pool was
                                  full (their node_count still advances, so the caller learns the size that
                                  would have sufficed), [2] scratch of dsb_node_append_rows                 */
    int32_t *page_fill;        /* [capacity / page_entries]: entries written in each claimed page (fused
                                  writer; NULL for a row-written store)                         */
    int32_t *node_count;       /* [n] in/out: nodes recorded per trajectory                     */
} dsb_node_store;

typedef struct {
    int64_t n;                 /* trajectories in this shard                                   */
    double *y;                 /* [dim][n] in/out                                              */
    double *t;                 /* [n] in/out                                                   */
    double *dt;                /* [n] in/out: step size to try next (reference self.dt)        */
    const double *params[DSB_MAX_PARAMS];   /* per-RHS parameter arrays (device)              */
    int64_t param_stride[DSB_MAX_PARAMS];   /* 1 = per trajectory, 0 = one shared value        */
    double tf, rtol, atol;
    int32_t *accepted;         /* [n] += accepted steps                                        */
    int32_t *rejected;         /* [n] += rejected attempts                                     */
    int32_t *status;           /* [n] dsb_traj_status (TOL_FAIL entries are skipped on entry)  */
    int64_t max_steps;         /* accepted-step budget per trajectory for this call, <0 = none */
    int32_t first_call;        /* 1: apply integrate()'s entry logic (:956-957, :971-974)      */
    dsb_node_store store;      /* dense output / step history (store.pool = NULL: off): one node per accepted
                                  step, plus node 0 = the initial condition of a trajectory that has none yet  */
    uint64_t *counters;        /* device, >= 8 words, zeroed by the caller: [0] work queue; on return
                                  [2] += step attempts of this call, [3] += accepted steps,
                                  [4] += trajectories left in DSB_TRAJ_TOL_FAIL, [5] += in DSB_TRAJ_RUNNING  */
    /* ---- ABI 2 ---- */
    int64_t stride;            /* leading dimension of y / y_init in elements (0: n).  A caller that streams
                                  an ensemble through the device in chunks passes pointers offset into one
                                  [dim][stride] allocation and n = the chunk length                          */
    const double *y_init;      /* [dim][stride] initial state read INSTEAD of y (NULL: y is in/out); y is
                                  then write-only, so re-running from the same y0 needs no reset copy      */
    double t_init, dt_init;    /* with fresh != 0: every trajectory starts at (t_init, dt_init)            */
    int32_t fresh;             /* != 0: t, dt, accepted, rejected, status are uninitialised outputs: the
                                  kernel starts from (t_init, dt_init) with zero counts and overwrites them
                                  (reference: OdeSystem.__init__/reset, differential_system.py:520-528,900-910) */
    /* flow control of a launch that overlaps its own uploads / downloads (set by dsb_erk_run_host) */
    int32_t done_shift;        /* done[i >> done_shift] counts the stored trajectories of block i >> done_shift */
    const uint32_t *avail;     /* device word (NULL: off): the initial state of trajectories [0, *avail) is resident;
                                  the kernel waits, bounded, before it loads a trajectory beyond it; a wait that
                                  times out sets counters[6] = 1 and abandons the remaining trajectories       */
    uint32_t *done;            /* device words (NULL: off), zeroed by the caller, see done_shift                  */
    /* ---- event detection on the device (n_events = 0: off) --------------------------------------------------
     * Replaces, per trajectory, prepare_events / handle_events and the event hooks of the integrate loop
     * (differential_system.py:29-57, 60-167, 1020-1066) for hyperplane event functions
     *     g_e(t, y) = sum_d w[e][d] * y[d] + wt[e] * t - c[e].
     * After every accepted step an event whose g changes sign over the step (fa * fb < 0, the bracketing test of
     * utilities/optimizer.py:252) is located as the root of g along the step's cubic Hermite interpolant
     * (utilities/interpolation.py:41-61; for a linear g that is a cubic in the step fraction), filtered by
     * direction, de-duplicated like :1043-1048, and recorded as (t_root, y(t_root), e).  A terminal event rolls the
     * step back and re-integrates to t_root exactly like the nested self.integrate(roots[-1]) of :1053-1054 (entry
     * clamp dt = |t_root - t| / 2, final-step clamp), then stops the trajectory with DSB_TRAJ_EVENT.            */
    int32_t n_events;          /* <= DSB_MAX_EVENTS                                                               */
    int32_t ev_capacity;       /* records per trajectory                                                          */
    const double *ev_coef;     /* device [n_events][dim + 2]: w[dim], wt, c                                       */
    const int32_t *ev_flags;   /* device [n_events]: bits 0-1 direction (0 either, 1 rising, 2 falling), bit 2 terminal */
    double *ev_records;        /* device [n][ev_capacity][dim + 2]: t, y[dim], event index                        */
    int32_t *ev_count;         /* device [n] in/out: events found (records beyond ev_capacity are dropped)        */
    const double *ev_coef_dy;  /* device [n_events][dim] weights on dState/dt (NULL: none): `requires_dstate` events,
                                  g(t, y, dy) = w.y + w_dy.dy + wt t - c with dy the interpolant's derivative
                                  (differential_system.py:92-96, utilities/interpolation.py:63-78)                */
    /* ---- ABI 3: scatter-store (out_y = NULL: off) ------------------------------------------------------------
     * The final gather of a sharded ensemble fused into the kernel's result store: the result of trajectory i goes
     * to index o = i * out_index_mul + out_index_add of the out_* arrays INSTEAD of index i of y, t, dt, accepted,
     * rejected, status (which are then not touched).  With out_index_mul = W ranks and out_index_add = rank the
     * round-robin shards land interleaved in one [dim][out_stride] buffer; the pointers may address another GPU's
     * memory (dsb_peer_open), the stores then travel over NVLink while the shard is still integrating.
     * Needs fresh != 0 and y_init, no dense output, no events; fast flavour, <= 7 stages. */
    double *out_y;             /* [dim][out_stride]                                                               */
    double *out_t, *out_dt;    /* [>= n * out_index_mul]                                                          */
    int32_t *out_accepted, *out_rejected, *out_status;
    int64_t out_stride, out_index_mul, out_index_add;
    /* ---- ABI 3: stop list (stops = NULL: off) -------------------------------------------------------------------
     * The reference serves `solve_ivp(..., t_eval=...)` by calling integrate(t) once per time (differential_system.py:
     * 1245-1248): every t_eval[k] is an exact stop of the integration.  With a stop list ONE launch does the same per
     * trajectory: integrate to stops[0] (entry logic :956-957, 971-974, final-step clamp :997-1002), record the state
     * in stop_y[0], continue to stops[1] with the entry logic re-applied and dt carried over (a final clamped step does
     * not update it, :1070-1071), ...  `tf` is ignored (the last stop is the final time); y, t, dt, accepted, rejected,
     * status receive the state at the last stop.  max_steps must be < 0; no events, node store or scatter-store. */
    const double *stops;       /* device [n_stops], monotone in the direction of integration                     */
    int32_t n_stops;
    double *stop_y;            /* device [n_stops][dim][stop_stride] out                                          */
    int64_t stop_stride;       /* >= n                                                                             */
} dsb_erk_run_args;

DSB_API int dsb_erk_run(dsb_erk_handle h, const dsb_erk_run_args *args, void *stream);

/* The same integration for an ensemble that lives in HOST memory (the reference keeps the state wherever y0
 * lives, differential_system.py:507-515): ONE persistent launch whose uploads and downloads overlap it.
 *   stream_up  : per chunk of 2^chunk_shift trajectories, cudaMemcpy2DAsync(host_y0 -> run.y) followed by a
 *                stream memory operation that publishes the resident count in flags[0];
 *   stream     : the kernel (run.avail = flags, run.done = flags + 1): a lane takes the next trajectory of the
 *                queue as soon as its initial state is resident, and counts stored results per chunk;
 *   stream_down: per chunk, a stream wait on flags[1 + c] == chunk length, then the device-to-host copies of the
 *                chunk's final state and counters.
 * `stream` finally waits for stream_down, so synchronising it means every result has landed.  No host
 * synchronisation inside.  `run` must describe device staging buffers with fresh = 1 and y_init = NULL; host
 * buffers should be pinned (pageable memory works but serialises).  Needs stream memory operations
 * (cuStreamWriteValue32 / cuStreamWaitValue32): DSB_ERR_UNSUPPORTED if the driver does not provide them -- the
 * caller then falls back to chunked dsb_erk_run launches. */
typedef struct {
    dsb_erk_run_args run;
    const double *host_y0;     /* [dim][host_stride]                                            */
    double *host_y;            /* [dim][host_stride] out                                        */
    int32_t *host_accepted, *host_rejected, *host_status;   /* [n] out; each may be NULL (not downloaded: the
                                  ensemble totals are in run.counters, the per-trajectory words stay readable in
                                  run.accepted / rejected / status).  In particular every status
                                  is DSB_TRAJ_DONE iff counters[4] + counters[5] == 0, so a caller can skip that
                                  download and fetch run.status only when a failure is reported               */
    int64_t host_stride;       /* leading dimension of host_y0 / host_y in elements (0: n)      */
    int32_t chunk_shift;       /* chunk = 2^chunk_shift trajectories                            */
    uint32_t *flags;           /* device, >= 1 + ceil(n / 2^chunk_shift) words (zeroed here)    */
    void *stream_up, *stream_down;
    /* ---- ABI 3 ---- */
    double *host_t;            /* [n] out, optional: final times                                */
    /* host_y0 == NULL: nothing is uploaded -- the initial state is resident (run.y_init, or run.y in place) and only the
     * result side of the pipeline runs; stream_up is unused.  The result pointers may then be any unified-address
     * destinations, e.g. another GPU's buffer mapped with dsb_peer_open: chunk c of the shard's results is copied to
     * the peer as soon as its trajectories are done, while the rest of the shard is still integrating. */
} dsb_erk_host_args;

DSB_API int dsb_erk_run_host(dsb_erk_handle h, const dsb_erk_host_args *args, void *stream);
/* number of kernel launches dsb_erk_run issues per call (for bench accounting) */
DSB_API int dsb_erk_run_launches(dsb_erk_handle h);

/* ---- stage-level explicit RK (arbitrary right-hand side between kernels) -----------------
 * For a right-hand side without a fused kernel (any torch callable; the dense linear RHS evaluated by
 * a GEMM) the host issues per attempt, on one stream and without synchronising:
 *   dsb_erk_stage_begin                       step start: final-step clamp, h, h0 (differential_system.py:997-1002;
 *                                             with first_call=1 also integrate()'s entry logic :956-957,:971-974)
 *   for s in stages: dsb_erk_stage_state(s)   y_stage = y + h*sum_j a_sj K_j, t_stage = t + c_s*h
 *                                             (integrators/components/runge_kutta_methods.py:44-54)
 *                    K[s] = rhs(t_stage, y_stage)          <- caller's code, same stream
 *   dsb_erk_stage_finish                      dY = h*sum b_j K_j, error estimate, per-trajectory L2 norm,
 *                                             controller, masked commit, counters
 *                                             (integrators/integrator_types.py:188-228,302-323,
 *                                              integrators/integrator_template.py:46-93,
 *                                              differential_system.py:1015-1018,1070-1071)
 * counters[1] holds the number of trajectories that still need attempts after dsb_erk_stage_finish.
 * Any explicit tableau with up to DSB_MAX_STAGES stages; every array is [dim][n], trajectory index fastest. */
#define DSB_MAX_STAGES 40

typedef struct {
    int64_t n;
    int32_t dim;               /* flattened state size of one trajectory                       */
    int32_t chunks;            /* partial-norm tiles over dim (>= 1)                           */
    double *y, *t, *dt;        /* committed state, as in dsb_erk_run_args                      */
    double *h;                 /* [n] step size of the attempt in flight                       */
    double *h0;                /* [n] first step size of the current retry chain               */
    const double *K[DSB_MAX_STAGES]; /* stage derivatives, each [dim][n]                       */
    double *y_stage;           /* [dim][n] out of dsb_erk_stage_state                          */
    double *t_stage;           /* [n]      out of dsb_erk_stage_state                          */
    double *dY;                /* [dim][n] scratch: increment of the attempt in flight         */
    double *scale;             /* [dim][n] controller scale (EMA on retries)                   */
    double *hist;              /* [3][n]  retry-chain history                                  */
    double *partial;           /* [chunks][n] partial squared error norms                      */
    int32_t *trial;            /* [n] attempt number inside the current step                   */
    int32_t *flags;            /* [n] bit0 active, bit1 final step, bit2 last attempt accepted */
    double tf, rtol, atol;
    int32_t *accepted, *rejected, *status;
    uint64_t *counters;        /* >= 4 words; [1] = active trajectories after finish           */
    int32_t one_step;          /* != 0: a trajectory goes inactive as soon as it accepts a step -- the
                                  integrator protocol `integrator(rhs, t, y, constants, timestep)`
                                  (integrators/integrator_template.py:38-40, differential_system.py:1003-1004) */
    /* event detection (n_events = 0: off): the hyperplane events of dsb_erk_run_args, for any right-hand side.
     * With events the attempt ends with dsb_erk_stage_finish_events(phase 0) -> f1 = rhs(t_stage, y_stage)
     * (the reference's final_rhs, integrators/integrator_types.py:308) -> dsb_erk_stage_finish_events(phase 1). */
    int32_t n_events, ev_capacity;
    const double *ev_coef;     /* [n_events][dim + 2]                                              */
    const int32_t *ev_flags;   /* [n_events]                                                       */
    double *ev_records;        /* [n][ev_capacity][dim + 2]                                        */
    int32_t *ev_count;         /* [n] in/out                                                       */
    double *ev_state;          /* [2 + 2*DSB_MAX_STAGE_EVENTS][n] scratch: t and dt before the last accepted step,
                                  last occurrence time per event, g at the step start per event     */
    double *tf_per;            /* [n] per-trajectory final time, initialised to tf by the caller; a terminal
                                  event re-targets it to the event time (NULL: tf for everyone)     */
    const double *ev_coef_dy;  /* [n_events][dim] weights on dState/dt (NULL: none): events of the reference's
                                  `requires_dstate` kind, g(t, y, dy) with dy = the interpolant's derivative
                                  (differential_system.py:92-96, utilities/interpolation.py:63-78)  */
    /* ABI 3: array-valued tolerances (integrators/integrator_types.py:37-38 takes `rtol` / `atol` of the state's
     * shape): one value per state component, shared by the trajectories; NULL = the scalar above                  */
    const double *rtol_vec, *atol_vec;   /* device [dim]                                          */
} dsb_erk_stage_args;

DSB_API int dsb_erk_stage_begin(dsb_erk_handle h, const dsb_erk_stage_args *a, int first_call, void *stream);
DSB_API int dsb_erk_stage_state(dsb_erk_handle h, const dsb_erk_stage_args *a, int stage, void *stream);
DSB_API int dsb_erk_stage_finish(dsb_erk_handle h, const dsb_erk_stage_args *a, void *stream);
/* finish with event detection: phase 0 = increment, error norm, controller, and y_stage = y + dY, t_stage = t_new
 * for the caller's f1 = rhs(t_stage, y_stage); phase 1 = locate / record events over the accepted steps (Hermite
 * interpolant from K[0] and f1), roll back and re-target trajectories stopped by a terminal event, masked commit. */
DSB_API int dsb_erk_stage_finish_events(dsb_erk_handle h, const dsb_erk_stage_args *a, int phase, const double *f1, void *stream);
/* Phase 1 with the event times located by the CALLER (any event callable, see dsb_event_search_*): roots is
 * [n_events][n], NaN = no event of that kind in the step just attempted.  Ordering in time, the repeat filter, the event
 * records, terminal roll-back / re-targeting and the masked commit are those of phase 1 above. */
DSB_API int dsb_erk_stage_finish_events_located(dsb_erk_handle h, const dsb_erk_stage_args *a, const double *f1,
                                                const double *roots, void *stream);
/* kernel launches issued by one begin + S x state + finish sequence (bench accounting) */
DSB_API int dsb_erk_stage_launches(dsb_erk_handle h);

/* ---- explicit symplectic splitting integrators -------------------------------------------
 * Replace ExplicitSymplecticIntegrator.__call__/step (integrators/integrator_types.py:375-417) and the
 * fixed-step loop of OdeSystem.integrate around it (differential_system.py:996-1002,1015-1018).
 * `tableau` is the class data: stages x 3, column 1 = drift weight (also the time advance), column 2 = kick
 * weight (integrators/explicit_integration_schemes.py:340-425).  Kick variables are rows dim0//2.. of axis 0 of
 * the state (integrator_types.py:359-362): `drift_count` = (dim0 // 2) * prod(dim[1:]) leading elements of the
 * flattened state are drift (position) variables, the rest kick (momentum) variables. */
typedef struct dsb_symp_plan *dsb_symp_handle;

DSB_API int dsb_symp_create(dsb_symp_handle *out, int device, int rhs_id, int dim, int drift_count, int stages,
                            const double *tableau /* stages x 3, host */, unsigned flags);
DSB_API int dsb_symp_destroy(dsb_symp_handle h);

/* fused: gravitational N-body, state per system (2, bodies, 3) = (q; v), one launch per integrate() */
typedef struct {
    int64_t n;                 /* systems                                                       */
    double *y;                 /* [2*bodies*3][n] in/out                                        */
    double *t, *dt;            /* [n] in/out                                                    */
    double tf;
    int32_t *steps;            /* [n] += steps taken                                            */
    int32_t *status;           /* [n] dsb_traj_status                                           */
    int64_t max_steps;         /* step budget per system for this call, < 0 = none              */
    int32_t first_call;
    int32_t bodies;
    const double *masses;      /* [bodies] device, shared by all systems                        */
    double G, eps2;            /* gravitational constant, Plummer softening squared             */
} dsb_symp_run_args;

DSB_API int dsb_symp_run(dsb_symp_handle h, const dsb_symp_run_args *args, void *stream);

/* stage-level, any right-hand side: per step  begin; for s in stages: f = rhs(t_stage, y_stage),
 * accumulate(s, f);  commit.  counters[1] = systems still short of tf after commit. */
typedef struct {
    int64_t n;
    int32_t dim;
    double *y, *t, *dt;
    double *h;                 /* [n] step size of the step in flight                           */
    double *dS;                /* [dim][n] accumulated increment                                */
    double *y_stage;           /* [dim][n] y + dS                                               */
    double *t_stage;           /* [n]                                                           */
    int32_t *flags, *steps, *status;
    double tf;
    uint64_t *counters;
    /* event detection (n_events = 0: off), as in dsb_erk_stage_args; the step then ends with
     * dsb_symp_stage_commit_events(phase 0) -> f1 = rhs(t_stage, y_stage) -> dsb_symp_stage_commit_events(phase 1) */
    int32_t n_events, ev_capacity;
    const double *ev_coef;
    const int32_t *ev_flags;
    double *ev_records;
    int32_t *ev_count;
    double *ev_state;          /* [2 + 2*DSB_MAX_STAGE_EVENTS][n] scratch                        */
    double *tf_per;            /* [n] per-system final time, initialised to tf by the caller     */
    int32_t *ev_scratch;       /* [n] scratch                                                    */
    const double *ev_coef_dy;  /* [n_events][dim] weights on dState/dt (NULL: none)              */
} dsb_symp_stage_args;

DSB_API int dsb_symp_stage_begin(dsb_symp_handle h, const dsb_symp_stage_args *a, int first_call, void *stream);
DSB_API int dsb_symp_stage_accumulate(dsb_symp_handle h, const dsb_symp_stage_args *a, int stage, const double *f, void *stream);
DSB_API int dsb_symp_stage_commit(dsb_symp_handle h, const dsb_symp_stage_args *a, void *stream);
/* commit with event detection: phase 0 advances the time (y_stage / t_stage already hold the state after the step, for
 * the caller's f1 = rhs(t_stage, y_stage)); phase 1 locates / records events on the step's Hermite interpolant
 * (f0 = the first stage's right-hand side = rhs at the step start), rolls back systems stopped by a terminal event
 * and commits the others. */
DSB_API int dsb_symp_stage_commit_events(dsb_symp_handle h, const dsb_symp_stage_args *a, int phase, const double *f0,
                                         const double *f1, void *stream);

DSB_API int dsb_symp_stage_commit_events_located(dsb_symp_handle h, const dsb_symp_stage_args *a, const double *f0,
                                                 const double *f1, const double *roots, void *stream);

/* ---- root search for arbitrary event functions ---------------------------------------------
 * Replaces the reference's vectorised Brent iteration on t -> event(t, sol(t)[, sol.grad(t)], **constants) over an
 * accepted step (differential_system.py:60-167, utilities/optimizer.py:178-366) for callables that are not affine
 * (those are solved in closed form inside the step kernels).  The callable is evaluated by the caller between the
 * kernels, batched over every (event, trajectory) pair; all arrays are [n_events][n] device memory:
 *   begin   bracket [0, 1] (step fractions) where g0 * g1 < 0 in an allowed direction for trajectories whose flag word
 *           has need_mask set and skip_mask clear; first iterate x and its time tq = t0 + x (t1 - t0);
 *           *counter (zeroed by the caller) = searches in flight
 *   step    Illinois update with fx = event(tq, y(tq), ...): new bracket, next x / tq; a converged search leaves the
 *           count.  The caller loops until the count is 0 or an iteration cap is reached.
 *   roots   event times, NaN where there is none -> dsb_erk_stage_finish_events_located */
typedef struct {
    int64_t n;
    int32_t n_events;
    double *lo, *hi, *flo, *fhi;   /* bracket in step fractions and the (Illinois-scaled) function values at its ends */
    double *x, *tq;                /* current iterate as a step fraction / as a time                                  */
    int32_t *side;                 /* -100 no event, 100 converged, else the end replaced last                        */
    const double *t0, *t1;         /* [n] step start / end times                                                      */
    uint64_t *counter;             /* device word                                                                     */
} dsb_event_search;

DSB_API int dsb_event_search_begin(const dsb_event_search *s, const double *g0, const double *g1, const int32_t *flags,
                                   int32_t need_mask, int32_t skip_mask, const int32_t *ev_flags, void *stream);
DSB_API int dsb_event_search_step(const dsb_event_search *s, const double *fx, void *stream);
DSB_API int dsb_event_search_roots(const dsb_event_search *s, double *roots, void *stream);

/* ---- dense linear right-hand side ---------------------------------------------------------
 * K = A @ Y for the linear benchmark system y' = A y over an ensemble stored as Y[n][columns]
 * (rhs_plugins.linear; the reference evaluates the user's `A @ y` inside compute_step,
 * integrators/components/runge_kutta_methods.py:49-54): C[M][N] = A[M][K] * B[K][N], row-major fp64, as a
 * hand-written FP64 tensor-core kernel (mma.sync.m8n8k4.f64, csrc/dgemm.cu).  Any M; N and K even and 16-byte aligned
 * pointers (rows are moved in 16-byte pieces), else DSB_ERR_UNSUPPORTED (the caller then uses a library GEMM).  Shapes
 * with M % 64 == 0, N % 128 == 0, K % 16 == 0 run the unpredicated instantiation. */
DSB_API int dsb_dgemm(const double *A, const double *B, double *C, int64_t M, int64_t N, int64_t K, void *stream);

/* The same product on the INT8 tensor cores (tcgen05.mma kind::i8, TMA, TMEM; csrc/ozaki_gemm.cu): every row of A and
 * every column of B is scaled by a power of two and cut into dsb_ozaki_slices() = 8 signed digit planes of 7 bits
 * (error-free: 55 bits below the largest element of the row / column); the 36 plane products of weight <= 7 are exact
 * INT8 x INT8 -> INT32 GEMMs, combined in FP64.  The result carries an error of about 2^-56 |row max| |column max| per
 * term plus one rounding -- tighter than a left-to-right FP64 dot product -- but it is NOT the rounding sequence of
 * dsb_dgemm / cuBLAS / numpy (tests compare it with an exactly rounded reference).
 *   dsb_ozaki_split_a   A[M][K] row-major -> planes[S][M][K] int8, scale[M]     (once per matrix; K % 16 == 0)
 *   dsb_ozaki_split_b   B[K][N] row-major -> planes[S][N][K] int8 (transposed), scale[N]; workspace: N * 8 bytes;
 *                       K % 128 == 0
 *   dsb_ozaki_gemm      C[M][N] = A B from the planes; M % 128 == 0, N % 64 == 0, K % 64 == 0, K <= 32768 (INT32
 *                       accumulators), else DSB_ERR_BAD_ARGUMENT; `sms` = CTAs to launch (one per SM); dbg_acc: NULL, or
 *                       [S][M][N] int32 receiving the raw accumulators (tests).  No hidden synchronisation. */
DSB_API int dsb_ozaki_slices(void);
DSB_API int dsb_ozaki_split_a(const double *A, int64_t M, int64_t K, signed char *planes, double *scale, void *stream);
DSB_API int dsb_ozaki_split_b(const double *B, int64_t K, int64_t N, signed char *planes, double *scale, void *workspace,
                              void *stream);
DSB_API int dsb_ozaki_gemm(const signed char *a_planes, const double *a_scale, const signed char *b_planes,
                           const double *b_scale, double *C, int64_t M, int64_t N, int64_t K, int sms, int32_t *dbg_acc,
                           void *stream);

/* ---- dense output / step history over the node store ---------------------------------------
 * Replace DenseOutput.__call__ / grad / find_interval (differential_system.py:205-246) and CubicHermiteInterp.__call__ /
 * grad (utilities/interpolation.py:41-78).  After a run the caller builds the index once:
 *   dsb_node_offsets   offsets[0..n] = exclusive prefix sum of node_count (offsets[n] = total nodes)
 *   dsb_node_locate    loc[offsets[i] + k] = entry of node k of trajectory i, for the first `claimed_entries`
 *                      (= cursor[0]) entries of the pool; *bad counts inconsistent headers (must stay 0)
 * and then evaluates any number of query times per trajectory in one launch. */
DSB_API int64_t dsb_node_offsets_scratch(int64_t n);        /* int64 words of scratch dsb_node_offsets needs */
DSB_API int dsb_node_offsets(const int32_t *node_count, int64_t n, int64_t *offsets /* [n + 1] */, int64_t *scratch, void *stream);
DSB_API int dsb_node_locate(const dsb_node_store *st, int64_t claimed_entries, int64_t n, const int64_t *offsets,
                            int64_t *loc /* [offsets[n]] */, uint64_t *bad /* device, zeroed */, void *stream);
/* Writer for the stage-level path (lock-step attempts; store.width == n): claims the next row-group and appends the node
 * (t[i], y[:, i], f[:, i]) for every trajectory i whose flags[i] has `mask` set (flags == NULL: all).  With the flag word
 * of dsb_erk_stage_args and mask = 4 this records exactly the steps the last dsb_erk_stage_finish accepted; f is
 * rhs(t, y) evaluated by the caller (the reference's final_rhs, integrators/integrator_types.py:308); f == NULL records
 * (t, y) only (step history without dense output). */
DSB_API int dsb_node_append_rows(const dsb_node_store *st, int64_t n, const double *t, const double *y, const double *f,
                                 const int32_t *flags, int32_t mask, void *stream);
/* out[q][d][i] = cubic Hermite interpolant (derivative = 0) or its time derivative (1) of trajectory i at
 * tq[i * tq_stride_traj + q * tq_stride_query], q < nq: a scalar time is strides (0, 0), one time per trajectory (1, 0),
 * a shared time grid (0, 1) -- the T x N evaluation behind `sol(t_eval)` / solve_ivp(t_eval=...). */
DSB_API int dsb_node_eval(const dsb_node_store *st, const int64_t *offsets, const int64_t *loc, int64_t n, const double *tq,
                          int64_t tq_stride_traj, int64_t tq_stride_query, int32_t nq, double *out, int32_t derivative,
                          void *stream);
/* Nodes k0 .. k0 + count - 1 of one trajectory as dense rows: t_out[j], y_out[j][dim] (f_out[j][dim] if not NULL) -- the
 * reference's per-step history `a.t`, `a.y` of a single system (differential_system.py:1015-1018, 1163-1196). */
DSB_API int dsb_node_gather(const dsb_node_store *st, const int64_t *offsets, const int64_t *loc, int64_t trajectory,
                            int64_t k0, int64_t count, double *t_out, double *y_out, double *f_out, void *stream);
/* One explicit Hermite step (the integrator protocol's dense_output(), integrators/integrator_types.py:109-119):
 * out[d][i] from p0, p1, m0, m1 [dim][n], t0 / t1 [n * t_stride] (t_stride = 0: shared), tq [n * tq_stride]. */
DSB_API int dsb_hermite_eval(const double *t0, const double *t1, int64_t t_stride, const double *p0, const double *p1,
                             const double *m0, const double *m1, int64_t n, int32_t dim, const double *tq, int64_t tq_stride,
                             double *out, int32_t derivative, void *stream);

/* ---- collectives of a sharded ensemble (NCCL over NVLink 5 / NVSwitch) -----------------------
 * The reference has no multi-device path; the ensemble engine shards trajectories round-robin over the GPUs of
 * one box (trajectory i -> rank i mod W) with NO exchange inside a step.  The only
 * collectives are the termination / statistics all-reduce over the kernels' own counters
 * (dsb_erk_run_args.counters[2..5]) and the final gather of states and per-trajectory counters.  Both are issued
 * on the caller's stream, behind the kernel, with no host synchronisation in between.
 * NCCL is bound at run time (dlopen of libnccl.so.2): every call returns DSB_ERR_UNSUPPORTED without it.
 * A communicator is either created here from a ncclUniqueId that the host side broadcasts
 * (dsb_comm_get_unique_id on rank 0 -> any out-of-band channel -> dsb_comm_init_rank on every rank) or wraps an
 * existing ncclComm_t (dsb_comm_from_nccl; not destroyed with the handle). */
typedef struct dsb_comm *dsb_comm_handle;
#define DSB_COMM_ID_BYTES 128
#define DSB_IPC_HANDLE_BYTES 64

DSB_API int dsb_comm_available(void);          /* 1: NCCL could be bound                          */
DSB_API int dsb_comm_nccl_version(void);       /* ncclGetVersion code, 0 if unavailable           */
DSB_API int dsb_comm_get_unique_id(void *id_out /* DSB_COMM_ID_BYTES, host */);
DSB_API int dsb_comm_init_rank(dsb_comm_handle *out, int device, int nranks, int rank,
                               const void *id /* DSB_COMM_ID_BYTES, host */);
DSB_API int dsb_comm_from_nccl(dsb_comm_handle *out, void *nccl_comm /* ncclComm_t */, int device);
DSB_API int dsb_comm_destroy(dsb_comm_handle c);
DSB_API int dsb_comm_size(dsb_comm_handle c);
DSB_API int dsb_comm_rank(dsb_comm_handle c);
/* recv[k] = sum over ranks of send[k] (send == recv allowed): the global termination test / ensemble statistics
 * straight from the device counters of dsb_erk_run -- no read-back in front of the collective. */
DSB_API int dsb_comm_allreduce_counters(dsb_comm_handle c, const uint64_t *send, uint64_t *recv, int count, void *stream);
/* Inverse of the round-robin sharding: full[r][j * W + w] = shard_w[r][j] on EVERY rank, for `rows` rows of
 * 4- or 8-byte elements (y: rows = dim; t, accepted, ...: rows = 1).  One in-place ncclAllGather into `staging`
 * (>= W * rows * ceil(n_global / W) * elem_bytes bytes, device) followed by one interleave kernel. */
DSB_API int dsb_comm_gather(dsb_comm_handle c, const void *shard, void *staging, void *full, int rows, int elem_bytes,
                            int64_t n_local, int64_t n_global, void *stream);
/* The interleave step alone, for shards that reached `staged` by other means (peer copies, dsb_erk_run_host with peer
 * destinations): full[r][j * W + w] = staged[w][r][j], staged = [W][rows][n_max] with n_max = ceil(n_global / W). */
DSB_API int dsb_interleave_shards(const void *staged, void *full, int rows, int elem_bytes, int64_t n_global, int world,
                                  void *stream);

/* Peer-visible device buffers (CUDA IPC) for the scatter-store of dsb_erk_run (out_* fields): the owner allocates
 * and exports, every other process of the box opens the handle and passes the mapped pointer to its own launches,
 * whose result stores then travel over NVLink while the integration is still running. */
DSB_API int dsb_peer_alloc(void **dev_ptr, int64_t bytes, void *ipc_handle_out /* DSB_IPC_HANDLE_BYTES */, int device);
DSB_API int dsb_peer_free(void *dev_ptr, int device);
DSB_API int dsb_peer_open(void **dev_ptr, const void *ipc_handle, int device);
DSB_API int dsb_peer_close(void *dev_ptr, int device);

#ifdef __cplusplus
}
#endif
#endif /* DESOLVER_B200_H */
