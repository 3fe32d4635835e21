/*
 * smm_b200.h -- C ABI of libsmm_b200.so, the B200-native (sm_100a) hot path of
 * LuzieH/SocialMediaModel: the stochastic ABM time-stepper, the mean-field PDE right-hand
 * side with fixed-step RK integration, and the ensemble drivers.
 *
 * The reference is pure Julia and has no FFI for this path; the boundary is its exported
 * Julia API (src/SocialMediaModel.jl:12-16).  Each entry point below is what a `ccall` in a
 * thin Julia shim binds in place of the Julia function it cites (see INTEGRATION.md and
 * socialmediamodel_b200/julia/SocialMediaModelB200.jl).  Plain pointers and sizes only.
 *
 * Conventions
 *  - Every array is laid out exactly like the Julia array named in its comment (column-major
 *    Float64 / Int32), so the shim passes `pointer(A)` without copying.  Written in C
 *    (row-major) index order the Julia array A[d1,d2,...,dk] is A[dk]...[d2][d1].
 *  - Indices crossing the ABI are 0-based (influencer 0..L-1); the reference's Float64
 *    1.0..L follower ids (abm.jl:164,212) are produced by the shim (+1).
 *  - Host entry points take HOST pointers, copy in/out themselves and return when the
 *    results are in the caller's buffers.  `*_dev` entry points take DEVICE pointers,
 *    enqueue on the context's stream and do not synchronise (use smm_ctx_sync).
 *  - The caller owns every buffer; the library keeps no pointer after return.
 *  - Every function returns 0 on success or a negative smm_status; the message of the last
 *    failure is available from smm_last_error().  No exception crosses the ABI and there is
 *    NO CPU fallback: without a CUDA device every compute entry point fails with
 *    SMM_ERR_CUDA.
 *  - smm_ctx_create binds a context to one GPU (one process per GPU: ensemble shards are selected by
 *    `first_realization` / the instance range and reduced by the caller, socialmediamodel_b200/dist.py);
 *    smm_ctx_create_multi puts several GPUs behind one context and shards inside the call.  Entry points
 *    are re-entrant per context, not thread-safe on the same context.
 */
#ifndef SMM_B200_H
#define SMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMM_VERSION 100 /* 0.1.0 */

typedef enum {
  SMM_OK = 0,
  SMM_ERR_INVALID = -1, /* bad argument (message says which) */
  SMM_ERR_CUDA = -2,    /* CUDA runtime/driver error, or no device */
  SMM_ERR_NOMEM = -3,   /* device or pinned-host allocation failed */
  SMM_ERR_UNSUPPORTED = -4
} smm_status;

/* arithmetic the state is held in on the device */
typedef enum {
  SMM_PREC_F32 = 0, /* throughput mode: FP32 positions/densities, FP64 switching decisions */
  SMM_PREC_F64 = 1  /* parity mode: everything Float64 like the reference */
} smm_precision;

typedef enum {
  SMM_NOISE_PHILOX = 0,  /* counter-based Philox4x32-10, key = seed, counter = (agent, step, stream, realization) */
  SMM_NOISE_INJECTED = 1 /* caller supplies every draw (parity against the reference/oracle) */
} smm_noise_mode;

/* fixed-step explicit Runge-Kutta schemes for the PDE (replaces DifferentialEquations.solve, pde.jl:291) */
typedef enum { SMM_RK_EULER = 0, SMM_RK_HEUN = 1, SMM_RK_SSPRK3 = 2, SMM_RK_RK4 = 3 } smm_scheme;

typedef struct smm_ctx smm_ctx; /* opaque: device, stream, scratch memory, CUDA graphs */

/* Model parameters `q` (setting.jl:1-22) plus the ABM simulation parameters of `p` (setting.jl:30-43). */
typedef struct {
  int32_t n, L, M; /* agents, influencers, media (M must be 2: abm.jl:12, pde.jl:204) */
  double eta, a, b, c, sigma, sigmahat, sigmatilde, frictionI, frictionM;
  double dt;                     /* ABMconstruct dt (setting.jl:32) */
  double dom_lo[2], dom_hi[2];   /* domain[d,1], domain[d,2] (setting.jl:35) */
  double controlspeed1, controltarget1[2], controlspeed2, controltarget2[2]; /* setting.jl:14-17 */
} smm_params;

/* Source of the random draws of the ABM (abm.jl:16,39,122,126,173,181-196). */
typedef struct {
  int32_t mode;         /* smm_noise_mode */
  uint64_t seed;        /* Philox key (the reference's chosenseed, abm.jl:155) */
  const double* xi;     /* injected N(0,1):  Julia xi[n,2,NT,R]  (abm.jl:173, one randn(n,2) per step) */
  const double* u1;     /* injected U[0,1):  Julia u1[n,NT,R]    (abm.jl:122) */
  const double* u2;     /* injected U[0,1):  Julia u2[n,NT,R]    (abm.jl:126; consumed only on a switch) */
  const double* xi_inf; /* injected N(0,1):  Julia [2,L,NT,R] or NULL = 0 (abm.jl:181,183) */
  const double* xi_med; /* injected N(0,1):  Julia [2,M,NT,R] or NULL = 0 (abm.jl:194,196) */
} smm_noise;

/* Regular grid: KDE grid of the ABM ensemble (setting.jl:37-38) and PDE grid (setting.jl:51-57). */
typedef struct {
  int32_t Nx, Ny;
  double x0, y0; /* domain[1,1], domain[2,1] */
  double dx, dy;
} smm_grid;

/* ---- library / context -------------------------------------------------------------- */
int smm_version(void);
int smm_device_count(void);                         /* 0 when no CUDA device is visible */
int smm_ctx_create(int device, smm_ctx** out);      /* binds to `device`; creates its stream */
/* Several GPUs of one box behind ONE context (single process).  ngpus <= 0: every visible device;
 * device_ids NULL: devices 0..ngpus-1.  The ensemble entry points called on such a context (smm_abm_ensemble,
 * smm_abm_ensemble_f32, smm_pde_solve) shard their realizations / instances over the GPUs -- contiguous blocks, one
 * host thread per GPU, every GPU writing its slice of the caller's arrays -- and reduce the ensemble sums with one
 * NCCL all-reduce (ncclCommInitAll communicators; libnccl is resolved with dlopen, host reduction without it).  All
 * other entry points run on the first device.  This is what lets the unchanged reference call
 * ABMsolveensemble(200, 1024) (ensemble.jl:11) use the whole box from one Julia task. */
int smm_ctx_create_multi(int ngpus, const int* device_ids, smm_ctx** out);
int smm_ctx_gpu_count(const smm_ctx* ctx);          /* GPUs behind the context (1 for smm_ctx_create) */
int smm_ctx_nccl_active(const smm_ctx* ctx);        /* 1 when the ensemble sums of a multi-GPU context go through NCCL */
int smm_ctx_destroy(smm_ctx* ctx);
int smm_ctx_sync(smm_ctx* ctx);                     /* wait for the context's stream */
int smm_ctx_set_stream(smm_ctx* ctx, void* cuda_stream); /* borrow a caller stream (e.g. torch's) */
const char* smm_last_error(const smm_ctx* ctx);     /* ctx may be NULL: last error of ctx_create */
/* pinned host memory, so that Julia/NumPy output arrays can be DMA targets */
int smm_host_alloc(smm_ctx* ctx, uint64_t bytes, void** out);
int smm_host_free(smm_ctx* ctx, void* p);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t smm_ctx_launch_count(const smm_ctx* ctx);
/* elapsed device time (ms, CUDA events on the context's stream) of the `which`-th kernel
 * class of the most recent call: 0 ABM step, 1 KDE, 2 ensemble reduce, 3 PDE convolution,
 * 4 PDE stencil/RK stage, 5 PDE moments.  Negative when that class did not run. */
double smm_ctx_kernel_ms(const smm_ctx* ctx, int which);

/* Measured denominators for the rooflines MEASURED_PEAKS.json does not carry, and for the floor of the end-to-end
 * ensemble call (bench.py): each runs a few-line kernel / copy, timed with CUDA events after a warm-up.
 *   SMM_MB_MUFU   result = G MUFU operations/s; size selects the operation: 0 ex2.approx, 1 sqrt.approx, 2 rsqrt.approx,
 *                 3 sqrt then ex2 (the dense ABM pair loop's chain), 4 rsqrt then ex2, 5 lg2, 6 rcp, 7 sin
 *   SMM_MB_GATHER result = G random 8-byte gathers/s from an L2-resident table of `size` entries (C5 neighbour reads)
 *   SMM_MB_LDS    result = GB/s of conflict-free 16-byte shared-memory loads (bound of the FFT passes)
 *   SMM_MB_D2H    result = GB/s of a `size`-byte device -> pinned-host copy
 *   SMM_MB_WIDEN  result = GB/s of Float64 written by the context's host threads widening `size` bytes of Float32 */
typedef enum { SMM_MB_MUFU = 0, SMM_MB_GATHER = 1, SMM_MB_LDS = 2, SMM_MB_D2H = 3, SMM_MB_WIDEN = 4 } smm_microbench_id;
int smm_microbench(smm_ctx* ctx, int32_t which, uint64_t size, double* result);

/* ---- ABM ---------------------------------------------------------------------------- */

/* ABMinit (abm.jl:8-45) on the device from the Philox INIT stream, for realizations
 * first_realization .. first_realization+R-1:  x ~ U(domain), state = +-1 w.p. 1/2,
 * quadrant followers, influencer = mean of its followers, media = [-1 -1; 1 1].
 * Outputs (host, any may be NULL): x0 Julia [n,2,R]; state [n,R]; follow0 [n,R] (0-based);
 * inf0 [L,2,R]; media0 [M,2,R].  L must be a square number (L = g*g cells of a g x g partition
 * of the domain, generalising abm.jl:21-24; L = 4 is the reference). */
int smm_abm_init(smm_ctx* ctx, const smm_params* q, uint64_t seed, int64_t first_realization, int32_t R,
                 double* x0, int32_t* state, int32_t* follow0, double* inf0, double* media0);

/* The draws the Philox mode uses for realization r, steps 0..NT-1 (for injecting the very same
 * numbers into the reference/oracle): xi Julia [n,2,NT], u1 [n,NT], u2 [n,NT]. */
int smm_abm_draws(smm_ctx* ctx, int32_t n, uint64_t seed, int64_t realization, int32_t NT,
                  double* xi, double* u1, double* u2);

/* ABMsolve (abm.jl:154-217) for R independent realizations sharing parameters and network.
 *   x0 [n,2,R], state [n,R], follow0 [n,R], inf0 [L,2,R], media0 [M,2,R]: initial condition
 *        (all NULL = on-device ABMinit from `noise->seed`, realizations first_realization..).
 *   rowptr/colidx: CSR follower network, row j = the i with IndNet[j,i]==1 (abm.jl:51; list j
 *        itself to keep the reference's self-loop, abm.jl:42); both NULL = all-to-all.
 *   savesteps[nsave]: ascending history indices in 0..NT (0 = initial state) to record.
 * Outputs (host; any may be NULL):
 *   xs      Julia [n,2,nsave,R]      positions            (abm.jl:209)
 *   follows Julia [n,nsave,R] int32  influencer followed, 0-based (abm.jl:212)
 *   infs    Julia [L,2,NT+1,R]       influencer history   (abm.jl:210)
 *   meds    Julia [M,2,NT+1,R]       media history        (abm.jl:211)
 *   numbers Julia [M,L,NT+1] int64   followers per (medium, influencer) and step, summed over
 *                                    the R realizations (plotting.jl:284-307) */
int smm_abm_solve(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                  int64_t first_realization,
                  const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0,
                  const double* media0, const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise,
                  int32_t nsave, const int32_t* savesteps,
                  double* xs, int32_t* follows, double* infs, double* meds, int64_t* numbers);

/* ABMsolveensemble (ensemble.jl:11-55): R realizations (global indices first_realization..),
 * KDE of the agents per (medium, influencer) on `grid` at the save steps.
 * Outputs (host; any may be NULL):
 *   us Julia [Nx,Ny,2,L,nsave,R], zs [2,2,nsave,R], ys [2,L,nsave,R]    (ensemble.jl:20-22,38-41)
 *   sum_us [Nx,Ny,2,L,nsave], sum_zs [2,2,nsave], sum_ys [2,L,nsave]: SUMS over the local R
 *        realizations -- the caller all-reduces over GPUs and divides by N (plotting.jl:361-363)
 *   numbers [M,L,NT+1] int64 as in smm_abm_solve. */
int smm_abm_ensemble(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                     int64_t first_realization,
                     const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0,
                     const double* media0, const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise,
                     int32_t nsave, const int32_t* savesteps, double kde_sigma, const smm_grid* grid,
                     double* us, double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys,
                     int64_t* numbers);

/* smm_abm_ensemble with the per-realization densities returned as Float32 (an extension outside the reference's
 * API for callers that accept it: `us` is 2.08 GB of Float64 for 1 024 realizations at 89x89, the FP32 state holds
 * no more than 24 bits of it).  us Julia Float32 [Nx,Ny,2,L,nsave,R]; everything else as in smm_abm_ensemble with
 * precision = SMM_PREC_F32.  Passing us = NULL to either entry point returns only the ensemble sums (means-only:
 * 1 MB instead of 2 GB). */
int smm_abm_ensemble_f32(smm_ctx* ctx, const smm_params* q, int32_t NT, int32_t R, int64_t first_realization,
                         const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0,
                         const double* media0, const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise,
                         int32_t nsave, const int32_t* savesteps, double kde_sigma, const smm_grid* grid,
                         float* us, double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys,
                         int64_t* numbers);

/* Device-resident variant of smm_abm_ensemble for callers that keep results on the GPU
 * (bench `value`, NCCL reduction of the sums by torch.distributed): Philox noise, on-device
 * init, all-to-all or CSR network already on the device.  Outputs are DEVICE pointers
 * (any may be NULL): us_dev Float32 [Nx,Ny,2,L,nsave,R] when precision is F32, Float64 when
 * F64; sum_us_dev Float64 [Nx,Ny,2,L,nsave]; sum_zs_dev [2,2,nsave]; sum_ys_dev [2,L,nsave];
 * numbers_dev int64 [M,L,NT+1].  Asynchronous on the context's stream. */
int smm_abm_ensemble_dev(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                         int64_t first_realization, uint64_t seed,
                         const int32_t* rowptr_dev, const int32_t* colidx_dev,
                         int32_t nsave, const int32_t* savesteps, double kde_sigma, const smm_grid* grid,
                         void* us_dev, double* sum_us_dev, double* sum_zs_dev, double* sum_ys_dev,
                         int64_t* numbers_dev);

/* Kernel-density estimate of one set of agents on `grid` (ensemble.jl:27-38, sumgaussian :2-8, gaussian
 * pde.jl:122): us Julia [Nx,Ny,2,L] = (1/n) sum over the agents with medium i and influencer j of the
 * Gaussian of width kde_sigma centred at the agent.  Also the density part of randominit (pde.jl:124-159).
 * x Julia [n,2]; follow [n] 0-based; state [n] in {-1,+1}. */
int smm_kde(smm_ctx* ctx, int32_t n, int32_t L, int32_t precision, const double* x, const int32_t* follow,
            const int32_t* state, double kde_sigma, const smm_grid* grid, double* us);

/* Large single population kept on the device between calls (config C5: n = 1e7, CSR network).
 * smm_abm_pop_create uploads the initial condition and network once; smm_abm_pop_step advances
 * `nsteps` steps with Philox noise (one fused kernel per step, no host round trip);
 * smm_abm_pop_read copies the current state back.  numbers [M,L] of the current state. */
typedef struct smm_abm_pop smm_abm_pop;
int smm_abm_pop_create(smm_ctx* ctx, const smm_params* q, int32_t precision, uint64_t seed,
                       const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0,
                       const double* media0, const int32_t* rowptr, const int32_t* colidx, smm_abm_pop** out);
int smm_abm_pop_step(smm_ctx* ctx, smm_abm_pop* pop, int32_t nsteps);
int smm_abm_pop_step_injected(smm_ctx* ctx, smm_abm_pop* pop, const double* xi /*[n,2]*/, const double* u1,
                              const double* u2);
int smm_abm_pop_read(smm_ctx* ctx, smm_abm_pop* pop, double* x /*[n,2]*/, int32_t* follow /*[n]*/,
                     double* inf /*[L,2]*/, double* media /*[M,2]*/, int64_t* numbers /*[M,L]*/);
/* steps taken so far, and how many switch decisions were re-evaluated in FP64 (guard band) */
int smm_abm_pop_stats(smm_ctx* ctx, smm_abm_pop* pop, int64_t* steps, int64_t* fp64_reevaluations);
int smm_abm_pop_destroy(smm_ctx* ctx, smm_abm_pop* pop);

/* ---- PDE ---------------------------------------------------------------------------- */

/* The right-hand side f(duzy, uzy, (p,q), t) (pde.jl:183-268) for B independent instances.
 *   q[B]: per-instance parameters (a,b,c,eta,sigma,frictionI/M, control fields; n,dt,dom unused)
 *   controlled_inf Julia [L,B] (0 free, 1 -> target1/speed1, 3 -> target2/speed2; pde.jl:239-255),
 *   controlled_med [2,B] (pde.jl:256-261); NULL = all free.
 *   u Julia [Nx,Ny,2,L,B], z [2,2,B], y [2,L,B]  ->  du, dz, dy (same shapes).
 * The interaction force F (pde.jl:49-59) is evaluated as a Toeplitz correlation generated from
 * w(|dix|,|diy|) on the fly; K_matrix/W_matrix (setting.jl:59) are never materialised. */
int smm_pde_rhs(smm_ctx* ctx, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L, int32_t precision,
                const double* controlled_inf, const double* controlled_med,
                const double* u, const double* z, const double* y, double* du, double* dz, double* dy);

/* PDEsolve / PDEsolveensemble (pde.jl:285-294, ensemble.jl:58-85) for B instances with a
 * fixed-step explicit RK scheme (the whole stage chain of one step is captured in a CUDA graph).
 *   savesteps[nsave]: ascending step indices in 0..nsteps whose state is recorded.
 * Outputs (host; any may be NULL): us Julia [Nx,Ny,2,L,nsave,B], zs [2,2,nsave,B], ys [2,L,nsave,B],
 *   sum_us [Nx,Ny,2,L,nsave], sum_zs, sum_ys: sums over the local B instances (plotting.jl:361-363). */
int smm_pde_solve(smm_ctx* ctx, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L, int32_t precision,
                  const double* controlled_inf, const double* controlled_med,
                  const double* u0, const double* z0, const double* y0,
                  double dt, int32_t nsteps, int32_t scheme, int32_t nsave, const int32_t* savesteps,
                  double* us, double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys);

/* Device-resident PDE batch kept between calls (bench `value`, parameter sweeps). */
typedef struct smm_pde_batch smm_pde_batch;
int smm_pde_batch_create(smm_ctx* ctx, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L,
                         int32_t precision, const double* controlled_inf, const double* controlled_med,
                         const double* u0, const double* z0, const double* y0, int32_t shared_init,
                         smm_pde_batch** out);
int smm_pde_batch_step(smm_ctx* ctx, smm_pde_batch* pb, double dt, int32_t nsteps, int32_t scheme);
int smm_pde_batch_read(smm_ctx* ctx, smm_pde_batch* pb, double* u, double* z, double* y);
int smm_pde_batch_sum_dev(smm_ctx* ctx, smm_pde_batch* pb, double* sum_u_dev, double* sum_z_dev, double* sum_y_dev);
int smm_pde_batch_destroy(smm_ctx* ctx, smm_pde_batch* pb);

/* Adaptive time integration with per-instance step-size control — the semantics of the reference's
 * DifferentialEquations.solve calls (pde.jl:291; controlstrategies.jl:10,96: abstol, reltol, dtmin,
 * force_dtmin, saveat) on an explicit embedded Runge-Kutta 3(2) pair (the 3-stage 3rd-order
 * Bogacki-Shampine scheme c = (0, 1/2, 3/4), b = (2/9, 1/3, 4/9), with the 2nd-order midpoint solution
 * b* = (0, 1, 0) as error estimator).  Every instance carries its own clock and step size on the device:
 * err = dt * sum_i (b_i - b*_i) k_i, norm = sqrt(mean((err / (abstol + reltol * max(|u_n|, |u_n+1|)))^2))
 * over all density values and ODE unknowns, accept if norm <= 1, next dt = dt * clamp(fac, 0.2, 5) with fac = 0.9 norm^(-7/30) prev^(2/15) after an accepted
 * step (PI control) and 0.9 norm^(-1/3) after a rejected one, no growth right after a rejection;
 * steps are shortened to land exactly on the requested output times. */
typedef struct {
  double abstol, reltol;   /* DifferentialEquations defaults: 1e-6, 1e-3 */
  double dt0;              /* first trial step; <= 0: 0.1 dx^2 / D */
  double dtmin;            /* smallest step (0 = none) */
  double dtmax;            /* largest step (<= 0 = none) */
  int32_t force_dtmin;     /* != 0: a step of size dtmin is accepted whatever its error (force_dtmin = true);
                              0: the call fails with SMM_ERR_UNSUPPORTED when the error test fails at dtmin */
  int32_t max_attempts;    /* per call and instance (<= 0: 1 000 000) */
} smm_adapt;

/* Advance every instance of the batch from the batch clock (0 at creation; smm_pde_batch_step advances it by
 * nsteps * dt) to t_end.  Optional outputs per instance [B]: accepted / rejected step counts of this call and
 * the step size the controller proposes next. */
int smm_pde_batch_advance(smm_ctx* ctx, smm_pde_batch* pb, double t_end, const smm_adapt* opt,
                          int64_t* naccept, int64_t* nreject, double* dt_next);

/* PDEsolve with the adaptive integrator: states recorded at savetimes[nsave] (ascending, >= 0; 0 = the
 * initial state, save_start = true).  Outputs as in smm_pde_solve; naccept / nreject [B] may be NULL. */
int smm_pde_solve_adaptive(smm_ctx* ctx, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L,
                           int32_t precision, const double* controlled_inf, const double* controlled_med,
                           const double* u0, const double* z0, const double* y0, const smm_adapt* opt,
                           int32_t nsave, const double* savetimes, double* us, double* zs, double* ys,
                           double* sum_us, double* sum_zs, double* sum_ys, int64_t* naccept, int64_t* nreject);

#ifdef __cplusplus
}
#endif
#endif /* SMM_B200_H */
