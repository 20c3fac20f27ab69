/* pof_b200 -- C ABI of the B200-native parallel-in-time IEKS hot path.
 *
 * The reference (nathanaelbosch/parallel-in-time-ode-filters, "pof") has no FFI of its own: the seam is cut at
 * its Python function signatures.  Each entry point below names the reference function it replaces.
 * Conventions: all pointers are DEVICE pointers unless stated "host"; fp64, row-major, contiguous, caller-owned;
 * nothing is allocated inside (the caller passes a workspace); every call is stream-ordered and does not
 * synchronise; return value 0 = ok, >0 = cudaError_t, <0 = argument error (POF_E_*).
 * No environment variables are read and there is no library-global mutable state: kernel-family choices are the
 * explicit `flags` argument, and the only resources a pass needs beyond the workspace (a side stream + two events
 * for the concurrent smoother up-sweep, optional timing events) live in a caller-owned context (pof_ctx_t).  Calls
 * with distinct workspaces and distinct contexts (or ctx = NULL) are re-entrant from different host threads.
 *
 * Shapes: N time points, n = N-1 steps, ODE dimension d, IWP order q, state dimension D = d*(q+1), state
 * ordering [y1, y1', .., y1^(q), y2, ..] (reference pof/transitions.py:80-88).
 */
#ifndef POF_B200_H
#define POF_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pof_stream_t; /* cudaStream_t */
typedef struct pof_ctx pof_ctx_t; /* opaque: side stream + events of one (concurrently running) pass */
typedef struct pof_p2p pof_p2p_t; /* opaque: peer-memory exchange area of one rank of a time-sharded run */
#define POF_P2P_HANDLE_BYTES 64   /* sizeof(cudaIpcMemHandle_t) */

/* flags (bitwise or) */
#define POF_F_FAMILY_TILE 1u    /* use the large-state kernels (one CTA per chunk / tree node) even where the
                                   register-resident d <= 4, D <= 16 family exists (cross-checks, tests) */
#define POF_F_TILE_SMEM_QR 2u   /* large-state kernels: shared-memory Householder sweeps instead of the register-
                                   resident ones (A/B measurement; slower on B200) */
#define POF_F_TREE_PER_LEVEL 4u /* one kernel launch per tree level instead of the dataflow sweeps (A/B measurement) */
#define POF_F_SMOOTH_TMA 8u     /* smoother: bulk-copy (TMA engine, cp.async.bulk + mbarrier) staging of the per-step
                                   backward kernels instead of plain global loads.  OFF by default: measured 30x
                                   SLOWER on B200 (64 independent 1 KB streams per SM) */
#define POF_F_TREE_UPDOWN 16u   /* filter tree: plain up-sweep / down-sweep over all levels instead of the hybrid whose
                                   top levels are ONE Kogge-Stone scan (half the dependent levels; A/B measurement) */

#define POF_E_UNSUPPORTED_DQ (-1) /* (d, q) combination not compiled in */
#define POF_E_WORKSPACE (-2)      /* workspace too small */
#define POF_E_ARG (-3)
#define POF_E_IVP (-4)            /* unknown built-in IVP id */

/* indices into the `scalars` output (device, POF_NSCALARS doubles) */
enum {
  POF_S_NLL = 0,        /* -sum log N(innovation)      reference filter.py:96-102                */
  POF_S_OBJ = 1,        /* smoother objective          reference smoother.py:20 (swapped args)   */
  POF_S_SSQ = 2,        /* sigma^2, reference formula  reference filter.py:105-114, utils.py:110 */
  POF_S_SSQ_PROPER = 3, /* sign-invariant sigma^2 = sum |S_L^{-1} y|^2 / (n d)                   */
  POF_S_NOT_CLOSE = 4,  /* #entries of means failing isclose(old,new,rtol=1e-13,atol=1e-8)
                           reference convergence_criteria.py:9                                    */
  POF_S_CSCALE = 5,     /* sqrt(sigma^2) applied to the returned chol (1 if !calibrate)           */
  POF_NSCALARS = 8
};

/* segments of a pass for the optional device timing of a context */
enum {
  POF_SEG_FOLD = 0, /* filter phase 1: chunk -> filtering element                                   */
  POF_SEG_FUP,      /* filter up-sweep incl. root (time-sharded stage A only)                        */
  POF_SEG_FTREE,    /* filter tree: up-sweep + down-sweep on one GPU, down-sweep in sharded stage B  */
  POF_SEG_SCAN,     /* filter phase 3: seeded square-root filter + backward kernels + statistics     */
  POF_SEG_SUP,      /* chunk smoothing elements + smoother up-sweep (side stream, concurrent to scan)*/
  POF_SEG_SDOWN,    /* smoother down-sweep                                                           */
  POF_SEG_SMOOTH,   /* smoother phase 3: seeded square-root RTS recursion, objective, calibration    */
  POF_SEG_COUNT
};

/* built-in vector fields (reference pof/ivp.py) for the fused f / Jacobian evaluation */
enum {
  POF_IVP_LOGISTIC = 0,      /* ivp.py:7-14    params: none                     */
  POF_IVP_LOTKAVOLTERRA = 1, /* ivp.py:17-31   params: a,b,c,d                  */
  POF_IVP_VANDERPOL = 2,     /* ivp.py:34-41   params: mu                       */
  POF_IVP_FITZHUGHNAGUMO = 3,/* ivp.py:44-60   params: a,b,tinv,l               */
  POF_IVP_ROBER = 4,         /* ivp.py:63-79   params: k1,k2,k3                 */
  POF_IVP_RIGIDBODY = 5,     /* ivp.py:82-90   params: p0,p1,p2                 */
  POF_IVP_SEIR = 6,          /* ivp.py:93-109  params: p0,p1,p2,p3              */
  POF_IVP_THREEBODY = 7,     /* ivp.py:112-126 params: mu                       */
  POF_IVP_HENONHEILES = 8,   /* ivp.py:137-152 params: p                        */
  POF_IVP_LORENZ96 = 9       /* synthetic larger-state problem (BASELINE config 5), params: forcing */
};

/* 1 if some kernel family serves (d, q); 1 if the large-state (CTA-per-chunk) family does (any d, 1 <= q <= 5,
 * D = d (q+1) limited by the 227 KB of shared memory per CTA: D <= 64 at d = 16) */
int pof_supported(int d, int q);
int pof_supported_tile(int d, int q);

/* Execution context (optional; NULL is valid everywhere and means: everything in line on the caller's stream).
 * With a context the smoother's up-sweep of a pass runs on the context's side stream concurrently with the filter
 * scan (fork / join through events: the pass stays stream-ordered on `s` and can be captured into a CUDA graph).
 * Create it on the device the passes run on; one context per concurrently running pass. */
int pof_ctx_create(pof_ctx_t** out);
void pof_ctx_destroy(pof_ctx_t* ctx);
/* measurement aid (bench.py): per-segment device time, CUDA events on the launching streams around each segment */
void pof_ctx_profile_enable(pof_ctx_t* ctx, int on);
int pof_ctx_profile_read(pof_ctx_t* ctx, double* ms_out /* POF_SEG_COUNT */, int64_t* count_out /* POF_SEG_COUNT */);

/* kernel-launch count of one pass; the device's FP64 FMA peak measured by a register-resident DFMA loop */
int64_t pof_launches_per_pass(int64_t N, int d, int q, int64_t chunk_len, uint32_t flags);
int pof_measure_dfma_tflops(pof_stream_t s, double* tflops_out /* host */);

/* default chunk length (steps per chunk) for a problem size: fills the `sm_count` SMs with one wave of chunks */
int64_t pof_default_chunk_len(int64_t N, int d, int q, int sm_count, uint32_t flags);

/* workspace bytes needed by pof_linear_filtsmooth_f64 / pof_ieks_* / pof_shard_* for the given chunk length */
size_t pof_workspace_bytes(int64_t N, int d, int q, int64_t chunk_len);

/* Batched associative operators on packed elements -- the reference's
 *   sqrt_filtering_operator(elem1, elem2)   pof/parallel_filtsmooth/filter.py:117-142
 *   sqrt_smoothing_operator(elem1, elem2)   pof/parallel_filtsmooth/smoother.py:53-63
 * Packed layouts per element (doubles): filter [A D*D | b D | U D*D | eta D | Z D*D], smoother [g D | E D*D | D D*D].
 * e1, e2, out: (n, elem) arrays.  elem1 = earlier in time for the filter; for the smoother the reference calls the
 * operator on the reversed sequence, so elem1 = LATER in time.  One warp per element pair. */
int pof_filter_combine_f64(pof_stream_t s, int64_t n, int D, const double* e1, const double* e2, double* out,
                           uint32_t flags);
int pof_smooth_combine_f64(pof_stream_t s, int64_t n, int D, const double* e1, const double* e2, double* out,
                           uint32_t flags);

/* Fused linearisation for the built-in IVPs -- replaces vmap(linearize)(om, states[1:])
 *   pof/step.py:12-22, pof/observations.py:35-40 with om = E1 x - f(E0 x) (pof/convenience.py:26-28):
 *   H_k = E1 - J_f(E0 m_{k+1}) E0,  c_k = J_f y - f(y), y = E0 m_{k+1},  k = 0..n-1.
 * scale0, scale1: the Nordsieck scalings E0 = scale0 * e_0^T, E1 = scale1 * e_1^T per block
 * (pof/transitions.py:53-68).  params: host pointer.  means_t1: the (n,D) rows of states t = 1..n (i.e. means[1:])
 * -> H (n,d,D), c (n,d). */
int pof_linearize_ivp_f64(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d, int q,
                          double scale0, double scale1, const double* means_t1, double* H, double* c);

/* One linear filter+smoother pass -- replaces
 *   pof.parallel_filtsmooth.linear_filtsmooth(x0, dtm, dom)   pof/parallel_filtsmooth/__init__.py:5-10
 * for the preconditioned IWP transition model (F = I_d (x) flip(pascal), QL = I_d (x) qL; pof/transitions.py:37-50)
 * and noiseless affine observations (cholR = 0; pof/observations.py:35-40), plus the calibration of
 * pof/step.py:42-44 and the means test of pof/convergence_criteria.py:9.
 *   qL_host  : host pointer, (q+1)x(q+1) lower Cholesky factor of flip(hilbert(q+1))
 *   x0_mean (D), x0_chol (D,D): initial state            H (n,d,D), c (n,d): linearised observation models
 *   means (N,D): IN the previous trajectory means (compared for convergence), OUT the smoothed means
 *   chols (N,D,D) or NULL: OUT smoothed Cholesky factors (lower triangular), times sqrt(sigma^2) if calibrate
 *   fmeans (N,D), fchols (N,D,D) or NULL: OUT filtered states (fchols are square-root factors, not triangular)
 *   scalars: OUT POF_NSCALARS doubles */
int pof_linear_filtsmooth_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t N, int d, int q,
                              int64_t chunk_len, const double* qL_host,
                              const double* x0_mean, const double* x0_chol, const double* H, const double* c,
                              double* means, double* chols, double* fmeans, double* fchols, int calibrate,
                              double* scalars, void* ws, size_t ws_bytes);

/* The same pass for a GENERAL linear-Gaussian model -- the full argument range of the reference's seam
 *   linear_filtsmooth(x0, TransitionModel(F (n,D,D), QL (n,D,D)), AffineModel(H (n,d,D), b (n,d), cholR (n,d,d)))
 *   pof/parallel_filtsmooth/__init__.py:5-10, pof/transitions.py:15-19, pof/observations.py:23-33:
 *   F, QL  : per-step dense transition matrices and lower-triangular process-noise factors, e.g. the reference's
 *            non-preconditioned models on a non-uniform grid (pof/transitions.py:71-99, pof/convenience.py:48-73);
 *            both NULL = the preconditioned IWP given by qL_host (then qL_host must not be NULL)
 *   cholR  : lower-triangular factors of the observation covariances (the reference's regularised iterations,
 *            pof/observations.py:43-83); NULL = noiseless
 * D = d (q+1).  Served by the large-state ("tile") kernels for any (d, q) they support (pof_supported_tile);
 * chunk_len from pof_default_chunk_len(.., POF_F_FAMILY_TILE), workspace from pof_workspace_bytes. */
int pof_linear_filtsmooth_general_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t N, int d, int q,
                                      int64_t chunk_len,
                                      const double* qL_host, const double* F, const double* QL, const double* x0_mean,
                                      const double* x0_chol, const double* H, const double* c, const double* cholR,
                                      double* means, double* chols, double* fmeans, double* fchols, int calibrate,
                                      double* scalars, void* ws, size_t ws_bytes);
/* One fused IEKS iteration for a built-in IVP -- replaces the body of the reference's while loop,
 *   pof.step.ieks_step(om, dtm, x0, states)   pof/step.py:33-45   (called from pof/solver.py:48-55):
 * linearise at means[1:] (fused f / Jacobian, kept in the compact form [J_f | c] inside the workspace: the dense
 * (n,d,D) H is never materialised), filter + smoother pass, calibration, convergence reductions.
 * means (N,D): IN previous trajectory, OUT smoothed means; chols (N,D,D) or NULL; scalars as above. */
int pof_ieks_iteration_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                           int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host, double scale0, double scale1,
                           const double* x0_mean, const double* x0_chol, double* means, double* chols, int calibrate,
                           double* scalars, void* ws, size_t ws_bytes);

/* The same iteration WITH the stopping rule evaluated on the device -- the body AND the condition of the reference's
 *   jax.lax.while_loop(cond, body, val)   pof/solver.py:36-57, pof/convergence_criteria.py:4-13.
 * loop_state (device, 8 doubles, zero-initialised by the caller): [0] stop flag, [1] iterations done, [2] obj and
 * [3] nll of the last iteration.  After the pass a one-thread kernel sets the flag when the rule fires
 * (isnan | isclose(obj_old, obj, rtol 1e-6, atol 1e-9) | no mean entry moved) or iterations > maxiters; once it is
 * set, every kernel of later calls returns at once and means / chols / scalars stay those of the final iteration.
 * The host may therefore enqueue several steps (or replays of a captured CUDA graph) without synchronising in between.
 * Register-resident family only (POF_E_UNSUPPORTED_DQ otherwise). */
int pof_ieks_loop_step_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                           int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                           double scale0, double scale1, const double* x0_mean, const double* x0_chol, double* means,
                           double* chols, int calibrate, double* scalars, double* loop_state, int64_t maxiters,
                           void* ws, size_t ws_bytes);

/* The WHOLE loop as one launch -- the reference's jax.lax.while_loop (pof/solver.py:36-57) as a CUDA graph: the body
 * (one pof_ieks_loop_step on fixed buffers) sits in a WHILE conditional node whose condition the stopping-rule kernel
 * sets on the device (cudaGraphSetConditional), so the loop runs to convergence / maxiters without the host.
 *   create : records the body for exactly the given buffers (all pointers are baked in and must stay valid; call
 *            after at least one eager pof_ieks_loop_step on the same workspace); no kernel runs.
 *   launch : enqueues the loop on stream s; stream-ordered, returns at once.  loop_state is NOT reset: zero it (or
 *            keep the count of earlier eager steps in it) before launching.  If its stop flag is already set the
 *            launch is a no-op.
 * Returns POF_E_UNSUPPORTED_DQ where pof_ieks_loop_step does, a CUDA error code if the driver cannot build
 * conditional nodes (callers then fall back to enqueuing pof_ieks_loop_step themselves). */
typedef struct pof_loop pof_loop_t;
int pof_ieks_loop_create_f64(pof_loop_t** out, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                             int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                             double scale0, double scale1, const double* x0_mean, const double* x0_chol,
                             double* means, double* chols, int calibrate, double* scalars, double* loop_state,
                             int64_t maxiters, void* ws, size_t ws_bytes);
int pof_ieks_loop_launch(pof_loop_t* loop, pof_stream_t s);
void pof_ieks_loop_destroy(pof_loop_t* loop);

/* Sequential extended Kalman smoother for a built-in IVP -- replaces
 *   pof.sequential_filtsmooth.filtsmooth(x0, dtm, om)   pof/sequential_filtsmooth/__init__.py:5-10
 * (EKF relinearised at the predicted mean of every step, filter.py:9-30, then RTS, smoother.py:8-28), the numerical
 * core of pof.solver.sequential_eks_solve (solver.py:76-96).  O(N) span by construction: one thread walks the grid
 * (baseline / cross-check path).  means (N,D), chols (N,D,D): OUT smoothed, uncalibrated.  scalars[POF_S_NLL] holds
 * +sum log-likelihood like the reference's sequential path (filter.py:91). */
int pof_sequential_eks_f64(pof_stream_t s, uint32_t flags, int ivp_id, const double* params_host, int nparams,
                           int64_t N, int d, int q,
                           const double* qL_host, double scale0, double scale1, const double* x0_mean,
                           const double* x0_chol, double* means, double* chols, double* scalars, void* ws,
                           size_t ws_bytes);

/* ---- time-sharded (multi-GPU) form of the same pass: three local stages with two exchange points ----------
 * Rank r owns the contiguous step range [k_lo, k_hi) of the global n steps; H, c are the LOCAL slices
 * (k_hi-k_lo steps) and means/chols/fmeans/fchols the LOCAL rows t in (k_lo, k_hi] (plus row t = 0 on rank 0:
 * local row index = t - t_lo with t_lo = k_lo + (k_lo > 0)), i.e. n_loc + (rank==0) rows.
 *  stage A: fold + up-sweep -> the rank's filtering element `carry_f` (3D^2+2D doubles)
 *           [exchange: all-gather carry_f; every rank folds the earlier ranks' elements onto x0 with
 *            pof_filter_apply_chain_f64 to get its incoming state]
 *  stage B: down-sweep from `state_in` (D + D*D), filter scan, smoother up-sweep -> `carry_s` (2D^2+D),
 *           `state_end` (D + D*D filtered state at k_hi), `partials` (3 doubles: nll, ssq_ref, ssq_proper sums)
 *           [exchange: all-gather carry_s and state_end; pof_smooth_apply_chain_f64 gives the rank's seed;
 *            all-reduce partials -> cscale]
 *  stage C: smoother down-sweep from `seed` (D + D*D smoothed state at k_hi), smoother scan
 *           -> means/chols, `partials2` (2 doubles: obj sum, not-close count); the objective term that couples
 *           the first local state to the previous rank's last state is added by the rank that owns the step. */
int pof_shard_stage_a_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host,
                          const double* H, const double* c, double* carry_f, void* ws, size_t ws_bytes);
int pof_shard_stage_b_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host,
                          const double* H, const double* c, const double* state_in, double* fmeans, double* fchols,
                          double* carry_s, double* state_end, double* partials, void* ws, size_t ws_bytes);
/* the same two stages reading the compact linearisation [J_f | c] (n_loc, d*d+d) written by
 * pof_linearize_ivp_compact_f64 instead of dense (H, c) */
int pof_linearize_ivp_compact_f64(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d,
                                  int q, double scale0, const double* means_t1, double* Jc);
int pof_shard_stage_a_compact_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host,
                                  const double* Jc, double scale0, double scale1, double* carry_f, void* ws,
                                  size_t ws_bytes);
int pof_shard_stage_b_compact_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host,
                                  const double* Jc, double scale0, double scale1, const double* state_in,
                                  double* fmeans, double* fchols, double* carry_s, double* state_end, double* partials,
                                  void* ws, size_t ws_bytes);
int pof_shard_stage_c_f64(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host,
                          const double* seed, int is_last_rank, int has_row0, const double* cscale, double* means,
                          double* chols, double* partials2, void* ws, size_t ws_bytes);
/* What a rank does with the all-gathered carries, ONE launch per exchange (register-resident family, D <= 16;
 * pof_shard_exchange_supported tells; otherwise use the chain entry points below):
 *  filter  : gathered = world payloads of `stride` doubles whose first 3D^2+2D doubles are carry_f of each rank;
 *            state_in (D + D*D) <- x0 combined with the carries of ranks 0 .. rank-1 (filter.py:117-142, state form)
 *  smoother: gathered payload of a rank = [carry_s (2D^2+D) | state_end (D+D*D) | partials (3)];
 *            sums the partials in rank order (bitwise identical on all ranks) -> scalars[NLL, SSQ, SSQ_PROPER,
 *            CSCALE] and *cscale = sqrt(sigma^2) (1 if !calibrate), sigma^2 = sum / n_steps_total / d;
 *            seed (D + D*D) <- the last rank's end state combined with the carries of ranks world-1 .. rank+1
 *  scalars : gathered = world pairs (obj, not-close count) -> scalars[OBJ], scalars[NOT_CLOSE]
 * scratch: D + D*D doubles. */
int pof_shard_exchange_supported(int D, uint32_t flags);
int pof_shard_exchange_filter_f64(pof_stream_t s, uint32_t flags, int D, int rank, int world, const double* gathered,
                                  int64_t stride, const double* x0_mean, const double* x0_chol, double* state_in,
                                  double* scratch);
int pof_shard_exchange_smooth_f64(pof_stream_t s, uint32_t flags, int D, int d, int rank, int world,
                                  int64_t n_steps_total, int calibrate, const double* gathered, int64_t stride,
                                  double* seed, double* scratch, double* cscale, double* scalars);
int pof_shard_exchange_scalars_f64(pof_stream_t s, int world, const double* gathered, double* scalars);

/* The same three exchanges over PEER MEMORY instead of a collective library (one process per GPU of one NVLink /
 * NVSwitch box, world <= 8): every rank owns an exchange area (slots for the payloads of all ranks, double-buffered,
 * plus flags); the exchange kernel stores this rank's payload into its slot of EVERY rank's area with plain stores
 * over NVLink, releases a flag there, waits for the flags of the ranks it needs in its own area and folds -- compute
 * and collective in ONE launch, no NCCL call, capturable into a CUDA graph.
 *   pof_p2p_create  : allocates the local area (cudaMalloc: CUDA IPC handles refer to whole allocations) and returns
 *                     its IPC handle; the caller exchanges the world handles (any host-side mechanism)
 *   pof_p2p_connect : opens the peers' areas (cudaIpcOpenMemHandle, enables peer access)
 *   pof_p2p_status  : 0, or 1 if an exchange gave up waiting for a peer after ~2 s (it never hangs the GPU)
 * payloads: filter = carry_f (3D^2+2D); smoother = [carry_s | state_end | partials (3) | pad] (2D^2+D + D+D*D + 4);
 * scalars = (obj, not-close). */
int pof_p2p_create(int rank, int world, int D, pof_p2p_t** out, unsigned char* handle_out /* POF_P2P_HANDLE_BYTES */);
int pof_p2p_connect(pof_p2p_t* p, const unsigned char* handles /* world x POF_P2P_HANDLE_BYTES, rank order */);
void pof_p2p_destroy(pof_p2p_t* p);
int pof_p2p_status(pof_p2p_t* p, int* status_host);
int pof_p2p_exchange_filter_f64(pof_stream_t s, uint32_t flags, pof_p2p_t* p, const double* carry_f,
                                const double* x0_mean, const double* x0_chol, double* state_in, double* scratch);
int pof_p2p_exchange_smooth_f64(pof_stream_t s, uint32_t flags, pof_p2p_t* p, int d, int64_t n_steps_total,
                                int calibrate, const double* payload, double* seed, double* scratch, double* cscale,
                                double* scalars);
int pof_p2p_exchange_scalars_f64(pof_stream_t s, pof_p2p_t* p, const double* pair, double* scalars);

/* state <- op(state, elems[0]), op(.., elems[1]), ... (count packed filter elements, earlier first) */
int pof_filter_apply_chain_f64(pof_stream_t s, uint32_t flags, int D, int count, const double* state_in,
                               const double* elems,
                               double* state_out, double* scratch /* >= D+D*D doubles */);
/* state <- smoothing op(state(later), elems[count-1]), ..., elems[0]  (elements in time order, applied last-first) */
int pof_smooth_apply_chain_f64(pof_stream_t s, uint32_t flags, int D, int count, const double* state_in,
                               const double* elems,
                               double* state_out, double* scratch);

/* Initial linearisation trajectory init="prior" (the default of pof.solver.solve) -- replaces
 *   pof.initialization.prior_init   pof/initialization.py:66-89  (get_initial_trajectory, convenience.py:88-90):
 * row 0 = x0 = (m0, 0); row k >= 1 = one prediction of x0 over the step size ts[k] (the ABSOLUTE time: the reference
 * passes steps = ts[1:]) with the non-preconditioned IWP model, in non-preconditioned coordinates:
 *   means[k] = P_k F PI_k m0,  chols[k] = tria([0, P_k QL]) = -P_k QL (LAPACK sign convention).
 * ts (N) device, m0 (D) device (the Taylor-mode initial mean, NOT preconditioned), qL_host as above;
 * means (N,D) out; chols (N,D,D) out or NULL (only the means feed the first linearisation). */
int pof_prior_init_f64(pof_stream_t s, int64_t N, int d, int q, const double* qL_host, const double* ts,
                       const double* m0, double* means, double* chols);

/* Final calibration + projection -- replaces pof/solver.py:66-71 (`chol *= sqrt(ssq)`; ys = E0 states):
 * ymean (N,d) = scale0 * means[:, b*(q+1)],  ychol (N,d,D) = mult * scale0 * chols[:, b*(q+1), :]. */
int pof_project_f64(pof_stream_t s, int64_t N, int d, int q, double scale0, const double* mult_dev,
                    const double* means, const double* chols, double* ymean, double* ychol);

/* ------------------------------------------------------------------------------------------------------------------
 * Optional fp32 mode (BASELINE north_star: "an optional fp32 mode reported separately"; the reference runs in fp32
 * unless JAX_ENABLE_X64 is set).  The register-resident kernel family (d <= 4, D <= 16) compiled with the scalar type
 * float: same entry points with the suffix _f32, same argument lists and semantics; every DEVICE array is float,
 * host-side arguments (qL_host, params_host, scalings) stay double, the workspace is sized by
 * pof_workspace_bytes_f32.  Not available in fp32: the large-state tile kernels (flags & POF_F_FAMILY_TILE, noisy
 * observations, general transition models, Lorenz-96), the sequential EKS.  Accuracy: the disagreement between
 * association orders grows like N^3; profiles/r02_fp32_accuracy.md lists where 1e-4 on the outputs holds. */
size_t pof_workspace_bytes_f32(int64_t N, int d, int q, int64_t chunk_len);
int pof_filter_combine_f32(pof_stream_t s, int64_t n, int D, const float* e1, const float* e2, float* out,
                           uint32_t flags);
int pof_smooth_combine_f32(pof_stream_t s, int64_t n, int D, const float* e1, const float* e2, float* out,
                           uint32_t flags);
int pof_linearize_ivp_f32(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d, int q,
                          double scale0, double scale1, const float* means_t1, float* H, float* c);
int pof_linearize_ivp_compact_f32(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d,
                                  int q, double scale0, const float* means_t1, float* Jc);
int pof_linear_filtsmooth_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t N, int d, int q,
                              int64_t chunk_len, const double* qL_host, const float* x0_mean, const float* x0_chol,
                              const float* H, const float* c, float* means, float* chols, float* fmeans, float* fchols,
                              int calibrate, float* scalars, void* ws, size_t ws_bytes);
int pof_ieks_iteration_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                           int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                           double scale0, double scale1, const float* x0_mean, const float* x0_chol, float* means,
                           float* chols, int calibrate, float* scalars, void* ws, size_t ws_bytes);
int pof_ieks_loop_step_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                           int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                           double scale0, double scale1, const float* x0_mean, const float* x0_chol, float* means,
                           float* chols, int calibrate, float* scalars, float* loop_state, int64_t maxiters, void* ws,
                           size_t ws_bytes);
int pof_ieks_loop_create_f32(pof_loop_t** out, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                             int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                             double scale0, double scale1, const float* x0_mean, const float* x0_chol, float* means,
                             float* chols, int calibrate, float* scalars, float* loop_state, int64_t maxiters,
                             void* ws, size_t ws_bytes);
int pof_shard_stage_a_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const float* H, const float* c, float* carry_f,
                          void* ws, size_t ws_bytes);
int pof_shard_stage_b_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const float* H, const float* c,
                          const float* state_in, float* fmeans, float* fchols, float* carry_s, float* state_end,
                          float* partials, void* ws, size_t ws_bytes);
int pof_shard_stage_a_compact_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host, const float* Jc, double scale0,
                                  double scale1, float* carry_f, void* ws, size_t ws_bytes);
int pof_shard_stage_b_compact_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host, const float* Jc, double scale0,
                                  double scale1, const float* state_in, float* fmeans, float* fchols, float* carry_s,
                                  float* state_end, float* partials, void* ws, size_t ws_bytes);
int pof_shard_stage_c_f32(pof_stream_t s, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const float* seed, int is_last_rank, int has_row0,
                          const float* cscale, float* means, float* chols, float* partials2, void* ws, size_t ws_bytes);
int pof_shard_exchange_filter_f32(pof_stream_t s, uint32_t flags, int D, int rank, int world, const float* gathered,
                                  int64_t stride, const float* x0_mean, const float* x0_chol, float* state_in,
                                  float* scratch);
int pof_shard_exchange_smooth_f32(pof_stream_t s, uint32_t flags, int D, int d, int rank, int world,
                                  int64_t n_steps_total, int calibrate, const float* gathered, int64_t stride,
                                  float* seed, float* scratch, float* cscale, float* scalars);
int pof_shard_exchange_scalars_f32(pof_stream_t s, int world, const float* gathered, float* scalars);
int pof_prior_init_f32(pof_stream_t s, int64_t N, int d, int q, const double* qL_host, const float* ts,
                       const float* m0, float* means, float* chols);
int pof_project_f32(pof_stream_t s, int64_t N, int d, int q, double scale0, const float* mult_dev, const float* means,
                    const float* chols, float* ymean, float* ychol);

#ifdef __cplusplus
}
#endif
#endif
