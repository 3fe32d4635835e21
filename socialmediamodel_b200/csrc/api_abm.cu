// api_abm.cu -- context management and the ABM entry points of the C ABI (include/smm_b200.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <algorithm>
#include <chrono>

#include <vector>
#include "smm_internal.h"
#include "host_widen.h"

std::string g_smm_last_error;

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int smm_version(void) { return SMM_VERSION; }

extern "C" int smm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int smm_ctx_create(int device, smm_ctx** out) {
  if (!out) return smm_fail(nullptr, SMM_ERR_INVALID, "smm_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return smm_fail(nullptr, SMM_ERR_CUDA, "smm_ctx_create: no CUDA device (%s); libsmm_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return smm_fail(nullptr, SMM_ERR_INVALID, "smm_ctx_create: device %d out of range [0,%d)", device, n);
  smm_ctx* ctx = new smm_ctx();
  ctx->device = device;
  for (int i = 0; i < SMM_NSCRATCH; i++) { ctx->scratch[i] = nullptr; ctx->scratch_bytes[i] = 0; }
  for (int i = 0; i < SMM_K_NCLASS; i++) { ctx->kernel_ms[i] = -1.0; ctx->ev_used[i] = false; ctx->ev0[i] = ctx->ev1[i] = nullptr; }
#define CREATE_TRY(expr)                                                                      \
  do {                                                                                        \
    cudaError_t e2 = (expr);                                                                  \
    if (e2 != cudaSuccess) {                                                                  \
      int rc = smm_fail(nullptr, SMM_ERR_CUDA, "smm_ctx_create: %s -> %s", #expr, cudaGetErrorString(e2)); \
      delete ctx;                                                                             \
      return rc;                                                                              \
    }                                                                                         \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    int rc = smm_fail(nullptr, SMM_ERR_UNSUPPORTED, "smm_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                      device, prop.major, prop.minor);
    delete ctx;
    return rc;
  }
  ctx->sm_count = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (int i = 0; i < SMM_K_NCLASS; i++) {
    CREATE_TRY(cudaEventCreate(&ctx->ev0[i]));
    CREATE_TRY(cudaEventCreate(&ctx->ev1[i]));
  }
#undef CREATE_TRY
  *out = ctx;
  return SMM_OK;
}

extern "C" int smm_ctx_destroy(smm_ctx* ctx) {
  if (!ctx) return SMM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SMM_NSCRATCH; i++)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < SMM_K_NCLASS; i++) {
    if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
    if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
  }
  for (int i = 0; i < 2; i++) { if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]); if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]); }
  for (int i = 0; i < 32; i++) if (ctx->ev_h[i]) cudaEventDestroy(ctx->ev_h[i]);
  if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
  smm_hostpool_destroy(ctx->hostpool);
  smm_multi_destroy(ctx->multi);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SMM_OK;
}

extern "C" int smm_ctx_sync(smm_ctx* ctx) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_ctx_sync: ctx is NULL");
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return smm_collect_times(ctx);
}

extern "C" int smm_ctx_set_stream(smm_ctx* ctx, void* cuda_stream) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_ctx_set_stream: ctx is NULL");
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return SMM_OK;
}

extern "C" const char* smm_last_error(const smm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_smm_last_error.c_str(); }

extern "C" int smm_host_alloc(smm_ctx* ctx, uint64_t bytes, void** out) {
  SMM_REQUIRE(ctx, ctx && out, "smm_host_alloc: NULL argument");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  SMM_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return SMM_OK;
}

extern "C" int smm_host_free(smm_ctx* ctx, void* p) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_host_free: ctx is NULL");
  if (p) SMM_CUDA(ctx, cudaFreeHost(p));
  return SMM_OK;
}

extern "C" int64_t smm_ctx_launch_count(const smm_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" double smm_ctx_kernel_ms(const smm_ctx* ctx, int which) {
  return (ctx && which >= 0 && which < SMM_K_NCLASS) ? ctx->kernel_ms[which] : -1.0;
}

int smm_scratch(smm_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) {
      SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      SMM_CUDA(ctx, cudaFree(ctx->scratch[slot]));
      ctx->scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
    }
    size_t cap = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return smm_fail(ctx, SMM_ERR_NOMEM, "device allocation of %zu bytes (scratch %d) failed: %s", cap, slot, cudaGetErrorString(e));
    }
    ctx->scratch_bytes[slot] = cap;
  }
  *out = ctx->scratch[slot];
  return SMM_OK;
}

int smm_pinned(smm_ctx* ctx, size_t bytes, void** out) {
  if (ctx->pinned_bytes < bytes) {
    if (ctx->pinned) { SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); SMM_CUDA(ctx, cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
    SMM_CUDA(ctx, cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocDefault));
    ctx->pinned_bytes = bytes;
  }
  *out = ctx->pinned;
  return SMM_OK;
}

void smm_tic(smm_ctx* ctx, int cls) {
  if (!ctx->ev_used[cls]) { cudaEventRecord(ctx->ev0[cls], ctx->stream); }
}
void smm_toc(smm_ctx* ctx, int cls) {
  cudaEventRecord(ctx->ev1[cls], ctx->stream);
  ctx->ev_used[cls] = true;
}
int smm_collect_times(smm_ctx* ctx) {
  for (int i = 0; i < SMM_K_NCLASS; i++) {
    if (ctx->ev_used[i]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->ev0[i], ctx->ev1[i]) == cudaSuccess) ctx->kernel_ms[i] = ms;
      else cudaGetLastError();
      ctx->ev_used[i] = false;
    }
  }
  return SMM_OK;
}
static void reset_times(smm_ctx* ctx) {
  for (int i = 0; i < SMM_K_NCLASS; i++) { ctx->kernel_ms[i] = -1.0; ctx->ev_used[i] = false; }
}

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------
static int check_params(smm_ctx* ctx, const smm_params* q, const char* who) {
  SMM_REQUIRE(ctx, q != nullptr, "%s: params is NULL", who);
  SMM_REQUIRE(ctx, q->n >= 1, "%s: n=%d must be >= 1", who, q->n);
  SMM_REQUIRE(ctx, q->L >= 1 && q->L <= 256, "%s: L=%d must be in 1..256", who, q->L);
  SMM_REQUIRE(ctx, q->M == 2, "%s: M=%d; the reference hard-wires M = 2 (abm.jl:12, pde.jl:204)", who, q->M);
  SMM_REQUIRE(ctx, q->dt > 0, "%s: dt must be > 0", who);
  SMM_REQUIRE(ctx, q->dom_hi[0] > q->dom_lo[0] && q->dom_hi[1] > q->dom_lo[1], "%s: empty domain", who);
  return SMM_OK;
}

static int upload(smm_ctx* ctx, int slot, const void* host, size_t bytes, void** dev) {
  SMM_TRY(smm_scratch(ctx, slot, bytes, dev));
  if (bytes) SMM_CUDA(ctx, cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SMM_OK;
}

// upload Float64 host data and convert to `real` on the device
static int upload_real(smm_ctx* ctx, int slot, const double* host, size_t count, int precision, void** dev) {
  if (precision == SMM_PREC_F64) return upload(ctx, slot, host, count * 8, dev);
  void* tmp;
  SMM_TRY(upload(ctx, S_STAGE, host, count * 8, &tmp));
  SMM_TRY(smm_scratch(ctx, slot, count * 4, dev));
  return smm_launch_convert_f64_to_real(ctx, (const double*)tmp, *dev, count, precision);
}

// device `real` -> Float64 host
static int download_real(smm_ctx* ctx, const void* dev, size_t count, int precision, double* host) {
  if (!count) return SMM_OK;
  const void* src = dev;
  if (precision != SMM_PREC_F64) {
    void* tmp;
    SMM_TRY(smm_scratch(ctx, S_STAGE, count * 8, &tmp));
    SMM_TRY(smm_launch_convert_real_to_f64(ctx, dev, (double*)tmp, count, precision));
    src = tmp;
  }
  SMM_CUDA(ctx, cudaMemcpyAsync(host, src, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return SMM_OK;
}

static int download(smm_ctx* ctx, const void* dev, size_t bytes, void* host) {
  if (bytes) SMM_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return SMM_OK;
}

static int isqrt_exact(int L) {
  int g = 1;
  while (g * g < L) g++;
  return g * g == L ? g : 0;
}

// device-side ABMinit into scratch
static int device_init(smm_ctx* ctx, const smm_params* q, int precision, uint64_t seed, int64_t first, int R,
                       void** x0, signed char** state, int32_t** follow0, double** inf0, double** media0) {
  const size_t rs = precision == SMM_PREC_F64 ? 8 : 4;
  const size_t n = q->n;
  int g = isqrt_exact(q->L);
  SMM_REQUIRE(ctx, g > 0, "ABMinit: L=%d is not a square number (L = g*g cells; the reference has L=4, abm.jl:13)", q->L);
  SMM_REQUIRE(ctx, first >= 0 && first + R <= 0xffffffffLL, "realization index out of the 32-bit counter range");
  SMM_TRY(smm_scratch(ctx, S_X0, (size_t)R * 2 * n * rs, x0));
  SMM_TRY(smm_scratch(ctx, S_STATE, (size_t)R * n, (void**)state));
  SMM_TRY(smm_scratch(ctx, S_FOLLOW, (size_t)R * n * 4, (void**)follow0));
  SMM_TRY(smm_scratch(ctx, S_INF0, (size_t)R * 2 * q->L * 8, (void**)inf0));
  SMM_TRY(smm_scratch(ctx, S_MED0, (size_t)R * 2 * q->M * 8, (void**)media0));
  AbmInitArgs I;
  I.k = make_abm_const(q);
  I.cells = g;
  I.seed = seed;
  I.first_realization = first;
  I.x0 = *x0; I.state = *state; I.follow0 = *follow0; I.inf0 = *inf0; I.media0 = *media0;
  return smm_launch_abm_init(ctx, I, R, precision);
}

extern "C" int smm_abm_init(smm_ctx* ctx, const smm_params* q, uint64_t seed, int64_t first_realization, int32_t R,
                            double* x0, int32_t* state, int32_t* follow0, double* inf0, double* media0) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_init: ctx is NULL");
  SMM_TRY(check_params(ctx, q, "smm_abm_init"));
  SMM_REQUIRE(ctx, R >= 1, "smm_abm_init: R must be >= 1");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  void* dx0; signed char* dst; int32_t* dfo; double* dinf; double* dmed;
  // Float64 so the caller sees exactly the positions the F64 kernels start from; the F32
  // kernels round the same Float64 values once
  SMM_TRY(device_init(ctx, q, SMM_PREC_F64, seed, first_realization, R, &dx0, &dst, &dfo, &dinf, &dmed));
  const size_t n = q->n;
  if (x0) SMM_TRY(download(ctx, dx0, (size_t)R * 2 * n * 8, x0));
  if (follow0) SMM_TRY(download(ctx, dfo, (size_t)R * n * 4, follow0));
  if (inf0) SMM_TRY(download(ctx, dinf, (size_t)R * 2 * q->L * 8, inf0));
  if (media0) SMM_TRY(download(ctx, dmed, (size_t)R * 2 * q->M * 8, media0));
  std::vector<signed char> st8;
  if (state) { st8.resize((size_t)R * n); SMM_TRY(download(ctx, dst, (size_t)R * n, st8.data())); }
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (state) for (size_t i = 0; i < st8.size(); i++) state[i] = st8[i];
  return SMM_OK;
}

extern "C" int smm_abm_draws(smm_ctx* ctx, int32_t n, uint64_t seed, int64_t realization, int32_t NT,
                             double* xi, double* u1, double* u2) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_draws: ctx is NULL");
  SMM_REQUIRE(ctx, n >= 1 && NT >= 1 && xi && u1 && u2, "smm_abm_draws: bad argument");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  void *dxi, *du1, *du2;
  size_t tot = (size_t)n * NT;
  SMM_TRY(smm_scratch(ctx, S_XI, tot * 2 * 8, &dxi));
  SMM_TRY(smm_scratch(ctx, S_U1, tot * 8, &du1));
  SMM_TRY(smm_scratch(ctx, S_U2, tot * 8, &du2));
  SMM_TRY(smm_launch_abm_draws(ctx, n, seed, (uint32_t)realization, NT, (double*)dxi, (double*)du1, (double*)du2));
  SMM_TRY(download(ctx, dxi, tot * 2 * 8, xi));
  SMM_TRY(download(ctx, du1, tot * 8, u1));
  SMM_TRY(download(ctx, du2, tot * 8, u2));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SMM_OK;
}

// ---------------------------------------------------------------------------------------
// ABM ensemble: shared implementation of smm_abm_solve / smm_abm_ensemble / _dev
// ---------------------------------------------------------------------------------------
namespace {

struct EnsembleRequest {
  const smm_params* q;
  int precision, NT, R;
  int64_t first;
  // host initial condition (all null = device init)
  const double* x0; const int32_t* state; const int32_t* follow0; const double* inf0; const double* media0;
  // network: host (rowptr/colidx) or device (rowptr_dev/colidx_dev)
  const int32_t* rowptr; const int32_t* colidx; const int32_t* rowptr_dev; const int32_t* colidx_dev;
  smm_noise noise;
  int nsave; const int32_t* savesteps;
  bool want_xs, want_follows, want_infs, want_meds, want_yz, want_numbers;
  // KDE
  bool want_kde; double kde_sigma; smm_grid grid;
  // device results
  void* d_xs; int32_t* d_follows; double* d_infs; double* d_meds; double* d_ys; double* d_zs; int64_t* d_numbers;
  signed char* d_state;
  void* d_us;  // caller-provided (dev API) or scratch
  double* ys_ext; double* zs_ext;  // caller-provided device slices (chunked ensembles) or null = scratch
  bool numbers_keep;               // accumulate into d_numbers instead of clearing it
  bool no_reset;                   // keep the kernel timers of the previous chunk
};

int run_ensemble(smm_ctx* ctx, EnsembleRequest& E, const char* who) {
  const smm_params* q = E.q;
  SMM_TRY(check_params(ctx, q, who));
  SMM_REQUIRE(ctx, E.precision == SMM_PREC_F32 || E.precision == SMM_PREC_F64, "%s: unknown precision %d", who, E.precision);
  SMM_REQUIRE(ctx, E.NT >= 0 && E.R >= 1, "%s: need NT >= 0 and R >= 1", who);
  SMM_REQUIRE(ctx, E.nsave >= 0 && (E.nsave == 0 || E.savesteps), "%s: savesteps missing", who);
  for (int m = 0; m < E.nsave; m++) {
    SMM_REQUIRE(ctx, E.savesteps[m] >= 0 && E.savesteps[m] <= E.NT, "%s: savesteps[%d]=%d outside 0..NT", who, m, E.savesteps[m]);
    SMM_REQUIRE(ctx, m == 0 || E.savesteps[m] > E.savesteps[m - 1], "%s: savesteps must be strictly ascending", who);
  }
  const bool host_init = E.x0 || E.state || E.follow0 || E.inf0 || E.media0;
  if (host_init)
    SMM_REQUIRE(ctx, E.x0 && E.state && E.follow0 && E.inf0 && E.media0, "%s: give all of x0,state,follow0,inf0,media0 or none", who);
  SMM_REQUIRE(ctx, (E.rowptr == nullptr) == (E.colidx == nullptr), "%s: rowptr and colidx must both be given or both be NULL", who);
  if (E.noise.mode == SMM_NOISE_INJECTED)
    SMM_REQUIRE(ctx, E.NT == 0 || (E.noise.xi && E.noise.u1 && E.noise.u2), "%s: injected noise needs xi, u1 and u2", who);
  else
    SMM_REQUIRE(ctx, E.noise.mode == SMM_NOISE_PHILOX, "%s: unknown noise mode %d", who, E.noise.mode);
  SMM_REQUIRE(ctx, E.first >= 0 && E.first + E.R <= 0xffffffffLL, "%s: realization index out of the 32-bit counter range", who);
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!E.no_reset) reset_times(ctx);

  const int precision = E.precision;
  const size_t rs = precision == SMM_PREC_F64 ? 8 : 4;
  const size_t n = q->n, L = q->L, M = q->M, R = E.R, NT = E.NT, nsave = E.nsave;

  AbmDenseArgs A;
  memset(&A, 0, sizeof A);
  A.k = make_abm_const(q);
  A.NT = E.NT; A.R = E.R; A.first_realization = E.first;

  // ---- initial condition
  void* dx0; signed char* dstate; int32_t* dfollow; double* dinf0; double* dmed0;
  if (host_init) {
    for (size_t i = 0; i < R * n; i++) {
      SMM_REQUIRE(ctx, E.follow0[i] >= 0 && E.follow0[i] < (int)L, "%s: follow0[%zu]=%d outside 0..L-1", who, i, E.follow0[i]);
      SMM_REQUIRE(ctx, E.state[i] == 1 || E.state[i] == -1, "%s: state[%zu]=%d is not +-1", who, i, E.state[i]);
    }
    SMM_TRY(upload_real(ctx, S_X0, E.x0, R * 2 * n, precision, &dx0));
    void* st32;
    SMM_TRY(upload(ctx, S_TMP0, E.state, R * n * 4, &st32));
    SMM_TRY(smm_scratch(ctx, S_STATE, R * n, (void**)&dstate));
    SMM_TRY(smm_launch_convert_i32_to_i8(ctx, (const int32_t*)st32, dstate, R * n));
    SMM_TRY(upload(ctx, S_FOLLOW, E.follow0, R * n * 4, (void**)&dfollow));
    SMM_TRY(upload(ctx, S_INF0, E.inf0, R * 2 * L * 8, (void**)&dinf0));
    SMM_TRY(upload(ctx, S_MED0, E.media0, R * 2 * M * 8, (void**)&dmed0));
  } else {
    SMM_TRY(device_init(ctx, q, precision, E.noise.seed, E.first, E.R, &dx0, &dstate, &dfollow, &dinf0, &dmed0));
  }
  A.x0 = dx0; A.state = dstate; A.follow0 = dfollow; A.inf0 = dinf0; A.media0 = dmed0;
  E.d_state = dstate;

  // ---- network
  if (E.rowptr) {
    SMM_REQUIRE(ctx, E.rowptr[0] == 0, "%s: rowptr[0] must be 0", who);
    for (size_t j = 0; j < n; j++) SMM_REQUIRE(ctx, E.rowptr[j + 1] >= E.rowptr[j], "%s: rowptr not monotone at %zu", who, j);
    size_t nnz = E.rowptr[n];
    for (size_t e = 0; e < nnz; e++) SMM_REQUIRE(ctx, E.colidx[e] >= 0 && E.colidx[e] < (int)n, "%s: colidx[%zu]=%d outside 0..n-1", who, e, E.colidx[e]);
    void *rp, *ci;
    SMM_TRY(upload(ctx, S_ROWPTR, E.rowptr, (n + 1) * 4, &rp));
    SMM_TRY(upload(ctx, S_COLIDX, E.colidx, nnz * 4, &ci));
    A.rowptr = (const int32_t*)rp; A.colidx = (const int32_t*)ci;
  } else if (E.rowptr_dev) {
    A.rowptr = E.rowptr_dev; A.colidx = E.colidx_dev;
  }

  // ---- noise
  A.noise_mode = E.noise.mode;
  A.seed = E.noise.seed;
  if (E.noise.mode == SMM_NOISE_INJECTED && NT > 0) {
    void* p;
    SMM_TRY(upload_real(ctx, S_XI, E.noise.xi, R * NT * 2 * n, precision, &p)); A.xi = p;
    SMM_TRY(upload(ctx, S_U1, E.noise.u1, R * NT * n * 8, &p)); A.u1 = (const double*)p;
    SMM_TRY(upload(ctx, S_U2, E.noise.u2, R * NT * n * 8, &p)); A.u2 = (const double*)p;
    if (E.noise.xi_inf) { SMM_TRY(upload(ctx, S_XIINF, E.noise.xi_inf, R * NT * L * 2 * 8, &p)); A.xi_inf = (const double*)p; }
    if (E.noise.xi_med) { SMM_TRY(upload(ctx, S_XIMED, E.noise.xi_med, R * NT * M * 2 * 8, &p)); A.xi_med = (const double*)p; }
  }

  // ---- outputs
  void* p;
  A.nsave = E.nsave;
  if (nsave) { SMM_TRY(upload(ctx, S_SAVESTEPS, E.savesteps, nsave * 4, &p)); A.savesteps = (const int32_t*)p; }
  // KDE of the save steps inside the step kernel where it can (FP32 symmetric path): no snapshot round trip, no second kernel
  const bool fused_kde = E.want_kde && nsave && smm_abm_dense_fuses_kde(ctx, A, precision, E.grid.Nx, E.grid.Ny);
  if (fused_kde) {
    const size_t per_snap = (size_t)E.grid.Nx * E.grid.Ny * 2 * L;
    if (!E.d_us) SMM_TRY(smm_scratch(ctx, S_US, R * nsave * per_snap * rs, &E.d_us));
    A.kde_us = E.d_us; A.kde_Nx = E.grid.Nx; A.kde_Ny = E.grid.Ny;
    A.kde_gx0 = E.grid.x0; A.kde_gy0 = E.grid.y0; A.kde_dx = E.grid.dx; A.kde_dy = E.grid.dy; A.kde_sigma = E.kde_sigma;
  }
  const bool need_snap = E.want_xs || (E.want_kde && !fused_kde);
  if (need_snap && nsave) { SMM_TRY(smm_scratch(ctx, S_XS, R * nsave * 2 * n * rs, &p)); A.xs = p; }
  if ((E.want_follows || (E.want_kde && !fused_kde)) && nsave) { SMM_TRY(smm_scratch(ctx, S_FOLS, R * nsave * n * 4, &p)); A.follows = (int32_t*)p; }
  if (E.want_infs) { SMM_TRY(smm_scratch(ctx, S_INFS, R * (NT + 1) * 2 * L * 8, &p)); A.infs = (double*)p; }
  if (E.want_meds) { SMM_TRY(smm_scratch(ctx, S_MEDS, R * (NT + 1) * 2 * M * 8, &p)); A.meds = (double*)p; }
  if (E.want_yz && nsave) {
    if (E.ys_ext) { A.ys = E.ys_ext; A.zs = E.zs_ext; }
    else {
      SMM_TRY(smm_scratch(ctx, S_YS, R * nsave * L * 2 * 8, &p)); A.ys = (double*)p;
      SMM_TRY(smm_scratch(ctx, S_ZS, R * nsave * M * 2 * 8, &p)); A.zs = (double*)p;
    }
  }
  if (E.want_numbers) {
    if (!E.d_numbers) { SMM_TRY(smm_scratch(ctx, S_NUMBERS, (NT + 1) * L * M * 8, &p)); E.d_numbers = (int64_t*)p; }
    if (!E.numbers_keep) SMM_CUDA(ctx, cudaMemsetAsync(E.d_numbers, 0, (NT + 1) * L * M * 8, ctx->stream));
    A.numbers = E.d_numbers;
  }

  smm_tic(ctx, SMM_K_ABM);
  SMM_TRY(smm_launch_abm_dense(ctx, A, precision));
  smm_toc(ctx, SMM_K_ABM);

  E.d_xs = A.xs; E.d_follows = A.follows; E.d_infs = A.infs; E.d_meds = A.meds; E.d_ys = A.ys; E.d_zs = A.zs;

  if (E.want_kde && nsave && !fused_kde) {
    const size_t per_snap = (size_t)E.grid.Nx * E.grid.Ny * 2 * L;
    if (!E.d_us) SMM_TRY(smm_scratch(ctx, S_US, R * nsave * per_snap * rs, &E.d_us));
    KdeArgs Kd;
    Kd.n = q->n; Kd.L = q->L; Kd.S = (int)(R * nsave); Kd.Nx = E.grid.Nx; Kd.Ny = E.grid.Ny;
    Kd.gx0 = E.grid.x0; Kd.gy0 = E.grid.y0; Kd.dx = E.grid.dx; Kd.dy = E.grid.dy; Kd.sigma = E.kde_sigma;
    Kd.xs = A.xs; Kd.follows = A.follows; Kd.state = dstate; Kd.nsave = E.nsave; Kd.us = E.d_us;
    smm_tic(ctx, SMM_K_KDE);
    SMM_TRY(smm_launch_kde(ctx, Kd, precision));
    smm_toc(ctx, SMM_K_KDE);
  }
  return SMM_OK;
}

int check_grid(smm_ctx* ctx, const smm_grid* g, const char* who) {
  SMM_REQUIRE(ctx, g != nullptr, "%s: grid is NULL", who);
  SMM_REQUIRE(ctx, g->Nx >= 1 && g->Ny >= 1 && g->dx > 0 && g->dy > 0, "%s: bad grid %dx%d dx=%g dy=%g", who, g->Nx, g->Ny, g->dx, g->dy);
  return SMM_OK;
}

}  // namespace

extern "C" int smm_abm_solve(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                             int64_t first_realization, const double* x0, const int32_t* state, const int32_t* follow0,
                             const double* inf0, const double* media0, const int32_t* rowptr, const int32_t* colidx,
                             const smm_noise* noise, int32_t nsave, const int32_t* savesteps, double* xs,
                             int32_t* follows, double* infs, double* meds, int64_t* numbers) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_solve: ctx is NULL");
  SMM_REQUIRE(ctx, noise != nullptr, "smm_abm_solve: noise is NULL");
  EnsembleRequest E;
  memset(&E, 0, sizeof E);
  E.q = q; E.precision = precision; E.NT = NT; E.R = R; E.first = first_realization;
  E.x0 = x0; E.state = state; E.follow0 = follow0; E.inf0 = inf0; E.media0 = media0;
  E.rowptr = rowptr; E.colidx = colidx; E.noise = *noise; E.nsave = nsave; E.savesteps = savesteps;
  E.want_xs = xs != nullptr; E.want_follows = follows != nullptr; E.want_infs = infs != nullptr;
  E.want_meds = meds != nullptr; E.want_numbers = numbers != nullptr;
  SMM_TRY(run_ensemble(ctx, E, "smm_abm_solve"));
  const size_t n = q->n, L = q->L, M = q->M;
  if (xs) SMM_TRY(download_real(ctx, E.d_xs, (size_t)R * nsave * 2 * n, precision, xs));
  if (follows) SMM_TRY(download(ctx, E.d_follows, (size_t)R * nsave * n * 4, follows));
  if (infs) SMM_TRY(download(ctx, E.d_infs, (size_t)R * (NT + 1) * 2 * L * 8, infs));
  if (meds) SMM_TRY(download(ctx, E.d_meds, (size_t)R * (NT + 1) * 2 * M * 8, meds));
  if (numbers) SMM_TRY(download(ctx, E.d_numbers, (size_t)(NT + 1) * L * M * 8, numbers));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return smm_collect_times(ctx);
}

// last-level cache of the host in bytes (0 = unknown): SMM_HOST_L3_MB, sysconf, then sysfs
static size_t host_l3_bytes() {
  if (const char* e = getenv("SMM_HOST_L3_MB")) return (size_t)std::max(0, atoi(e)) << 20;
#ifdef _SC_LEVEL3_CACHE_SIZE
  const long v = sysconf(_SC_LEVEL3_CACHE_SIZE);
  if (v > 0) return (size_t)v;
#endif
  if (FILE* f = fopen("/sys/devices/system/cpu/cpu0/cache/index3/size", "r")) {
    unsigned long kb = 0;
    char unit = 0;
    const int got = fscanf(f, "%lu%c", &kb, &unit);
    fclose(f);
    if (got >= 1 && kb > 0) return (size_t)kb * (unit == 'M' ? (size_t)1 << 20 : (size_t)1 << 10);
  }
  return 0;
}

// `us_any` is the caller's Float64 array, or (us_f32) a Float32 array of the same shape that receives the FP32
// device densities without widening (half the PCIe bytes, half the host bytes written).
static int abm_ensemble_impl(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                             int64_t first_realization, const double* x0, const int32_t* state,
                             const int32_t* follow0, const double* inf0, const double* media0,
                             const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                             const int32_t* savesteps, double kde_sigma, const smm_grid* grid, void* us_any, bool us_f32,
                             double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys,
                             int64_t* numbers) {
  double* us = us_f32 ? nullptr : static_cast<double*>(us_any);
  float* us32 = us_f32 ? static_cast<float*>(us_any) : nullptr;
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_ensemble: ctx is NULL");
  SMM_REQUIRE(ctx, !us32 || precision == SMM_PREC_F32, "smm_abm_ensemble_f32: Float32 densities need precision = SMM_PREC_F32");
  SMM_REQUIRE(ctx, noise != nullptr, "smm_abm_ensemble: noise is NULL");
  SMM_TRY(check_grid(ctx, grid, "smm_abm_ensemble"));
  SMM_REQUIRE(ctx, kde_sigma > 0, "smm_abm_ensemble: kde_sigma must be > 0");
  SMM_TRY(check_params(ctx, q, "smm_abm_ensemble"));
  SMM_REQUIRE(ctx, R >= 1 && NT >= 0 && nsave >= 0, "smm_abm_ensemble: need R >= 1, NT >= 0, nsave >= 0");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = q->n, L = q->L, M = q->M;
  const size_t per_real = (size_t)grid->Nx * grid->Ny * 2 * L * nsave;
  // The per-realization densities are by far the largest output (2 GB for 1 024 realizations at 89x89):
  // realizations are processed in chunks and the device->host copy of chunk c overlaps the
  // simulation of chunk c+1 (copy stream, double-buffered Float64 staging).
  int nchunks = 1;
  // Measured (1 024 realizations, Float64 `us`, 16 host threads; ms per call, tools/scratch/e2e_chunks.sh): 4 chunks x 2 copy
  // units (128 realizations per unit) 27.5-29.8; 4 x 4: 26.2; 8 x 2: 25.7; 8 x 3: 24.1; 8 x 4 (32 per unit): 23.4-24.7; 10 x 3:
  // 23.5; 12 x 2: 24.8 -- shorter lead-in before the first copy and finer interleaving of DMA and widening.  64 or more
  // units per call (8 x 8, 4 x 16, 16 x 4) fall off a cliff (68-113 ms): the unit count stays at 32.
  if (us_any && nsave && R >= 512) nchunks = std::min(8, R / 128);
  if (const char* ce = getenv("SMM_ENSEMBLE_CHUNKS")) { const int c = atoi(ce); if (c >= 1 && us_any && nsave && R >= 2 * c) nchunks = c; }
  const int Rc0 = (R + nchunks - 1) / nchunks;
  if (nchunks > 1 && !ctx->copy_stream) {
    SMM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      SMM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
      SMM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    }
  }
  void *d_ys_all = nullptr, *d_zs_all = nullptr, *d_numbers = nullptr, *d_sum_us = nullptr;
  const bool want_yz = zs || ys || sum_zs || sum_ys;
  if (want_yz && nsave) {
    SMM_TRY(smm_scratch(ctx, S_YS_ALL, (size_t)R * nsave * L * 2 * 8, &d_ys_all));
    SMM_TRY(smm_scratch(ctx, S_ZS_ALL, (size_t)R * nsave * M * 2 * 8, &d_zs_all));
  }
  if (numbers) SMM_TRY(smm_scratch(ctx, S_NUMBERS, (size_t)(NT + 1) * L * M * 8, &d_numbers));
  if (sum_us && nsave) SMM_TRY(smm_scratch(ctx, S_SUMUS, per_real * 8, &d_sum_us));
  // FP32 mode: the chunks cross PCIe as FP32 (half the bytes) into a pinned ring and host threads widen
  // them into `us` (host_widen.cpp); each chunk is copied in kSub units so that DMA and widening overlap.
  // Host widening halves the PCIe bytes but needs host cores and memory bandwidth (108 GB/s of Float64 output on a
  // 16-core host).  Measured per pass and rank: 1 GPU / 16 threads 28-35 ms with it, 41 ms without; 2 GPUs / 12
  // threads each 51 vs 65 ms; 4 GPUs / 8 threads each 99 vs 82 ms; 8 GPUs / 4 threads each 202 vs 181 ms.  Hence:
  // widen on the host when the rank has >= 12 threads (cores / LOCAL_WORLD_SIZE), else convert on the device and
  // let the DMA write the destination directly.  SMM_HOST_WIDEN=1 / 0 forces the choice.
  const char* hw_env = getenv("SMM_HOST_WIDEN");
  bool host_widen = nchunks > 1 && !us32 && precision == SMM_PREC_F32 && !(hw_env && hw_env[0] == '0');
  if (host_widen) {
    if (!ctx->hostpool) ctx->hostpool = smm_hostpool_create(ctx->host_threads);
    if (smm_hostpool_threads(ctx->hostpool) < 12 && !(hw_env && hw_env[0] == '1')) host_widen = false;
  }
  // Copy units and the pinned ring they land in, sized against the host's last-level cache (measured on a 16-core Xeon
  // with 60 MB of L3, 1 024 realizations, ms per call; profiles/r2c_e2e_chunk_sweep.txt):
  //  * ring smaller than the L3 (two slots of <= 22 realizations): the lines a DMA write targets are still cached from the
  //    widening threads' reads, every write has to snoop them out, and the copies crawl: 60-110 ms;
  //  * ring 1.1-1.4 x L3 (two slots of 32-43 realizations): 23.5-25 -- old lines are gone before they are rewritten, and the
  //    widening threads still find part of a freshly copied unit in the cache;
  //  * ring 2-4 x L3 (or 256 MB in 8-32 slots of any size): 26-30 -- everything the widening reads comes from DRAM.
  // Hence units of 0.55 x L3 in a ring of at least 1.3 x L3 (three slots of 32 realizations here: 24.2-24.5, the same as two
  // slots, with more distance to the first regime; on the slower boxes of the pool units of 32 gave 29.8 against 31-33 for
  // units of 43); with an unknown L3 a 256 MB ring (never in the first regime).
  // SMM_HOST_L3_MB, SMM_HOST_WIDEN_SUB (units per chunk) and SMM_HOST_RING_MB override.
  // The regimes were measured on a part with one large shared L3.  A small reported L3 usually means a cache per core
  // complex (several of them behind the widening threads, i.e. an aggregate the sizing cannot see): keep the 256 MB ring there.
  static const size_t l3_detected = host_l3_bytes();
  const size_t l3_bytes = l3_detected >= ((size_t)48 << 20) || getenv("SMM_HOST_L3_MB") ? l3_detected : 0;
  int kSub = 4;
  if (l3_bytes) {
    const size_t want = std::max<size_t>(1, (size_t)(0.55 * (double)l3_bytes) / (per_real * 4));
    kSub = (int)(((size_t)Rc0 + std::min<size_t>(want, (size_t)Rc0) - 1) / std::min<size_t>(want, (size_t)Rc0));
  }
  if (const char* e = getenv("SMM_HOST_WIDEN_SUB")) kSub = std::max(1, atoi(e));
  const size_t unit_real = ((size_t)Rc0 + kSub - 1) / kSub;  // realizations per copy unit
  int ring_slots = 2;
  if (host_widen) {
    const size_t unit_bytes = unit_real * per_real * 4;
    size_t ring_min = l3_bytes ? (size_t)(1.3 * (double)l3_bytes) : (size_t)256 << 20;
    if (const char* e = getenv("SMM_HOST_RING_MB")) ring_min = (size_t)std::max(0, atoi(e)) << 20;
    ring_slots = (int)std::min<size_t>(32, std::max<size_t>(2, (ring_min + unit_bytes - 1) / unit_bytes));
    for (int i = 0; i < ring_slots; i++)
      if (!ctx->ev_h[i]) SMM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_h[i], cudaEventDisableTiming));
    const size_t need = (size_t)ring_slots * unit_bytes;
    if (ctx->h_ring_bytes < need) {
      if (ctx->h_ring) { SMM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream)); SMM_CUDA(ctx, cudaFreeHost(ctx->h_ring)); ctx->h_ring = nullptr; ctx->h_ring_bytes = 0; }
      SMM_CUDA(ctx, cudaHostAlloc(&ctx->h_ring, need, cudaHostAllocDefault));
      ctx->h_ring_bytes = need;
    }
  }
  struct Pending { int chunk; size_t r0, rc; };
  Pending pending = {-1, 0, 0};
  void* stage[2] = {nullptr, nullptr};
  // copy units of the host-widen pipeline: unit g covers realizations [g_r0, g_r0 + g_rc) of chunk g / kSub
  struct Unit { int chunk; size_t r0, rc; bool last_of_chunk; };
  std::vector<Unit> units;
  size_t issued = 0, widened = 0;
  double dbg_sync = 0, dbg_widen = 0;
  auto issue_unit = [&](size_t g) -> int {
    const Unit& u = units[g];
    const int sl = u.chunk & 1, hs = (int)(g % (size_t)ring_slots);
    float* h = (float*)ctx->h_ring + (size_t)hs * unit_real * per_real;
    const float* d = (const float*)stage[sl] + (u.r0 - (size_t)u.chunk * Rc0) * per_real;
    SMM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[sl], 0));
    SMM_CUDA(ctx, cudaMemcpyAsync(h, d, u.rc * per_real * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    SMM_CUDA(ctx, cudaEventRecord(ctx->ev_h[hs], ctx->copy_stream));
    if (u.last_of_chunk) SMM_CUDA(ctx, cudaEventRecord(ctx->ev_copied[sl], ctx->copy_stream));
    return SMM_OK;
  };
  // widen every unit below `drain_end`; copies may run ahead into units below `issuable_end`
  auto pump = [&](size_t drain_end, size_t issuable_end) -> int {
    while (widened < drain_end) {
      while (issued < std::min(issuable_end, widened + 2)) SMM_TRY(issue_unit(issued++));
      const int hs = (int)(widened % (size_t)ring_slots);
      const auto t0 = std::chrono::steady_clock::now();
      SMM_CUDA(ctx, cudaEventSynchronize(ctx->ev_h[hs]));
      const auto t1 = std::chrono::steady_clock::now();
      const Unit& u = units[widened];
      smm_hostpool_widen(ctx->hostpool, (const float*)ctx->h_ring + (size_t)hs * unit_real * per_real,
                         us + u.r0 * per_real, u.rc * per_real);
      const auto t2 = std::chrono::steady_clock::now();
      dbg_sync += std::chrono::duration<double, std::milli>(t1 - t0).count();
      dbg_widen += std::chrono::duration<double, std::milli>(t2 - t1).count();
      widened++;
    }
    return SMM_OK;
  };
  auto issue_copy = [&](const Pending& pc) -> int {  // D2H of a staged chunk on the copy stream
    const int sl = pc.chunk & 1;
    SMM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[sl], 0));
    if (us32) SMM_CUDA(ctx, cudaMemcpyAsync(us32 + pc.r0 * per_real, stage[sl], pc.rc * per_real * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    else SMM_CUDA(ctx, cudaMemcpyAsync(us + pc.r0 * per_real, stage[sl], pc.rc * per_real * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    SMM_CUDA(ctx, cudaEventRecord(ctx->ev_copied[sl], ctx->copy_stream));
    return SMM_OK;
  };
  for (int c = 0; c < nchunks; c++) {
    const size_t r0 = (size_t)c * Rc0, rc = std::min((size_t)Rc0, (size_t)R - r0);
    if (rc == 0) break;
    EnsembleRequest E;
    memset(&E, 0, sizeof E);
    E.q = q; E.precision = precision; E.NT = NT; E.R = (int)rc; E.first = first_realization + (int64_t)r0;
    if (x0) { E.x0 = x0 + r0 * 2 * n; E.state = state ? state + r0 * n : nullptr; E.follow0 = follow0 ? follow0 + r0 * n : nullptr;
              E.inf0 = inf0 ? inf0 + r0 * 2 * L : nullptr; E.media0 = media0 ? media0 + r0 * 2 * M : nullptr; }
    else { E.x0 = nullptr; E.state = state; E.follow0 = follow0; E.inf0 = inf0; E.media0 = media0; }
    E.rowptr = rowptr; E.colidx = colidx; E.noise = *noise;
    if (noise->mode == SMM_NOISE_INJECTED) {
      if (noise->xi) E.noise.xi = noise->xi + r0 * NT * 2 * n;
      if (noise->u1) E.noise.u1 = noise->u1 + r0 * NT * n;
      if (noise->u2) E.noise.u2 = noise->u2 + r0 * NT * n;
      if (noise->xi_inf) E.noise.xi_inf = noise->xi_inf + r0 * NT * L * 2;
      if (noise->xi_med) E.noise.xi_med = noise->xi_med + r0 * NT * M * 2;
    }
    E.nsave = nsave; E.savesteps = savesteps;
    E.want_kde = us_any || sum_us; E.kde_sigma = kde_sigma; E.grid = *grid;
    E.want_yz = want_yz; E.want_numbers = numbers != nullptr;
    if (want_yz && nsave) { E.ys_ext = (double*)d_ys_all + r0 * nsave * L * 2; E.zs_ext = (double*)d_zs_all + r0 * nsave * M * 2; }
    E.d_numbers = (int64_t*)d_numbers; E.numbers_keep = c > 0; E.no_reset = c > 0;
    SMM_TRY(run_ensemble(ctx, E, "smm_abm_ensemble"));
    if (sum_us && nsave) {
      smm_tic(ctx, SMM_K_REDUCE);
      SMM_TRY(smm_launch_ensemble_sum(ctx, E.d_us, precision, (int)rc, per_real, (double*)d_sum_us, c > 0));
      smm_toc(ctx, SMM_K_REDUCE);
    }
    if (us_any && nsave) {
      if (nchunks == 1) {
        if (us32) SMM_TRY(download(ctx, E.d_us, rc * per_real * 4, us32));
        else SMM_TRY(download_real(ctx, E.d_us, rc * per_real, precision, us));
      } else if (host_widen) {
        const int sl = c & 1;
        SMM_TRY(smm_scratch(ctx, sl ? S_STAGE_B : S_STAGE_A, (size_t)Rc0 * per_real * 4, &stage[sl]));
        if (c >= 2) SMM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[sl], 0));  // staging slot free again
        SMM_CUDA(ctx, cudaMemcpyAsync(stage[sl], E.d_us, rc * per_real * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        SMM_CUDA(ctx, cudaEventRecord(ctx->ev_ready[sl], ctx->stream));
        const size_t drain_end = units.size();  // everything of the chunks before this one
        for (size_t o = 0; o < rc; o += unit_real)
          units.push_back({c, r0 + o, std::min(unit_real, rc - o), o + unit_real >= rc});
        // this chunk's kernels are queued: drain the previous chunk while they run
        SMM_TRY(pump(drain_end, units.size()));
      } else {
        const int sl = c & 1;
        SMM_TRY(smm_scratch(ctx, sl ? S_STAGE_B : S_STAGE_A, (size_t)Rc0 * per_real * (us32 ? 4 : 8), &stage[sl]));
        if (c >= 2) SMM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[sl], 0));  // staging slot free again
        if (us32) SMM_CUDA(ctx, cudaMemcpyAsync(stage[sl], E.d_us, rc * per_real * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        else SMM_TRY(smm_launch_convert_real_to_f64(ctx, E.d_us, (double*)stage[sl], rc * per_real, precision));
        SMM_CUDA(ctx, cudaEventRecord(ctx->ev_ready[sl], ctx->stream));
        // the previous chunk's copy is issued only now, after this chunk's kernels are queued, so that a
        // host-blocking copy into pageable memory still overlaps them
        if (pending.chunk >= 0) SMM_TRY(issue_copy(pending));
        pending = {c, r0, rc};
      }
    }
  }
  if (pending.chunk >= 0) SMM_TRY(issue_copy(pending));
  if (host_widen) SMM_TRY(pump(units.size(), units.size()));
  if (host_widen && getenv("SMM_E2E_DEBUG")) fprintf(stderr, "[smm] widen pipeline: %zu units, waiting for copies %.2f ms, widening %.2f ms\n", units.size(), dbg_sync, dbg_widen);
  void* p;
  // shard mode (member of a multi-GPU context): the sums stay in device scratch (S_SUMUS / S_SUMZS / S_SUMYS /
  // S_NUMBERS) for the NCCL all-reduce in multi.cu, which also delivers them to the caller
  const bool deliver = !ctx->shard_mode;
  if (sum_us && nsave && deliver) SMM_TRY(download(ctx, d_sum_us, per_real * 8, sum_us));
  if (sum_zs && nsave) {
    SMM_TRY(smm_scratch(ctx, S_SUMZS, nsave * M * 2 * 8, &p));
    SMM_TRY(smm_launch_small_sum(ctx, (const double*)d_zs_all, R, (int)(nsave * M * 2), (double*)p));
    if (deliver) SMM_TRY(download(ctx, p, nsave * M * 2 * 8, sum_zs));
  }
  if (sum_ys && nsave) {
    SMM_TRY(smm_scratch(ctx, S_SUMYS, nsave * L * 2 * 8, &p));
    SMM_TRY(smm_launch_small_sum(ctx, (const double*)d_ys_all, R, (int)(nsave * L * 2), (double*)p));
    if (deliver) SMM_TRY(download(ctx, p, nsave * L * 2 * 8, sum_ys));
  }
  if (zs && nsave) SMM_TRY(download(ctx, d_zs_all, (size_t)R * nsave * M * 2 * 8, zs));
  if (ys && nsave) SMM_TRY(download(ctx, d_ys_all, (size_t)R * nsave * L * 2 * 8, ys));
  if (numbers && deliver) SMM_TRY(download(ctx, d_numbers, (size_t)(NT + 1) * L * M * 8, numbers));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_stream) SMM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return smm_collect_times(ctx);
}

// An error return must not leave DMA into the caller's arrays (or into the pinned ring and the staging slots the
// next call reuses) in flight: drain both streams before handing the error back.
static int abm_ensemble_guarded(smm_ctx* ctx, int rc) {
  if (rc != SMM_OK && ctx) {
    const std::string keep = ctx->err;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaGetLastError();
    ctx->err = keep;
  }
  return rc;
}

extern "C" int smm_abm_ensemble(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                                int64_t first_realization, const double* x0, const int32_t* state,
                                const int32_t* follow0, const double* inf0, const double* media0,
                                const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                                const int32_t* savesteps, double kde_sigma, const smm_grid* grid, double* us,
                                double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys,
                                int64_t* numbers) {
  if (ctx && ctx->multi)
    return smm_multi_abm_ensemble(ctx, q, precision, NT, R, first_realization, x0, state, follow0, inf0, media0, rowptr, colidx, noise,
                                  nsave, savesteps, kde_sigma, grid, us, false, zs, ys, sum_us, sum_zs, sum_ys, numbers);
  return abm_ensemble_guarded(ctx, abm_ensemble_impl(ctx, q, precision, NT, R, first_realization, x0, state, follow0, inf0, media0, rowptr,
                                                     colidx, noise, nsave, savesteps, kde_sigma, grid, us, false, zs, ys, sum_us,
                                                     sum_zs, sum_ys, numbers));
}

extern "C" int smm_abm_ensemble_f32(smm_ctx* ctx, const smm_params* q, int32_t NT, int32_t R, int64_t first_realization,
                                    const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0,
                                    const double* media0, const int32_t* rowptr, const int32_t* colidx,
                                    const smm_noise* noise, int32_t nsave, const int32_t* savesteps, double kde_sigma,
                                    const smm_grid* grid, float* us, double* zs, double* ys, double* sum_us,
                                    double* sum_zs, double* sum_ys, int64_t* numbers) {
  if (ctx && ctx->multi)
    return smm_multi_abm_ensemble(ctx, q, SMM_PREC_F32, NT, R, first_realization, x0, state, follow0, inf0, media0, rowptr, colidx,
                                  noise, nsave, savesteps, kde_sigma, grid, us, true, zs, ys, sum_us, sum_zs, sum_ys, numbers);
  return abm_ensemble_guarded(ctx, abm_ensemble_impl(ctx, q, SMM_PREC_F32, NT, R, first_realization, x0, state, follow0, inf0, media0,
                                                     rowptr, colidx, noise, nsave, savesteps, kde_sigma, grid, us, true, zs, ys,
                                                     sum_us, sum_zs, sum_ys, numbers));
}

// entry for multi.cu: one shard of a multi-GPU ensemble on the member context `ctx`
int smm_abm_ensemble_shard(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R, int64_t first_realization,
                           const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0, const double* media0,
                           const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                           const int32_t* savesteps, double kde_sigma, const smm_grid* grid, void* us, bool us_f32, double* zs,
                           double* ys, double* sum_us, double* sum_zs, double* sum_ys, int64_t* numbers) {
  return abm_ensemble_guarded(ctx, abm_ensemble_impl(ctx, q, precision, NT, R, first_realization, x0, state, follow0, inf0, media0, rowptr,
                                                     colidx, noise, nsave, savesteps, kde_sigma, grid, us, us_f32, zs, ys, sum_us,
                                                     sum_zs, sum_ys, numbers));
}

extern "C" int smm_abm_ensemble_dev(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R,
                                    int64_t first_realization, uint64_t seed, const int32_t* rowptr_dev,
                                    const int32_t* colidx_dev, int32_t nsave, const int32_t* savesteps,
                                    double kde_sigma, const smm_grid* grid, void* us_dev, double* sum_us_dev,
                                    double* sum_zs_dev, double* sum_ys_dev, int64_t* numbers_dev) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_ensemble_dev: ctx is NULL");
  SMM_TRY(check_grid(ctx, grid, "smm_abm_ensemble_dev"));
  SMM_REQUIRE(ctx, kde_sigma > 0, "smm_abm_ensemble_dev: kde_sigma must be > 0");
  SMM_REQUIRE(ctx, (rowptr_dev == nullptr) == (colidx_dev == nullptr), "smm_abm_ensemble_dev: rowptr/colidx mismatch");
  EnsembleRequest E;
  memset(&E, 0, sizeof E);
  E.q = q; E.precision = precision; E.NT = NT; E.R = R; E.first = first_realization;
  E.rowptr_dev = rowptr_dev; E.colidx_dev = colidx_dev;
  E.noise.mode = SMM_NOISE_PHILOX; E.noise.seed = seed;
  E.nsave = nsave; E.savesteps = savesteps;
  E.want_kde = us_dev || sum_us_dev; E.kde_sigma = kde_sigma; E.grid = *grid; E.d_us = us_dev;
  E.want_yz = sum_zs_dev || sum_ys_dev; E.want_numbers = numbers_dev != nullptr; E.d_numbers = numbers_dev;
  SMM_TRY(run_ensemble(ctx, E, "smm_abm_ensemble_dev"));
  const size_t L = q->L, M = q->M;
  const size_t per_real = (size_t)grid->Nx * grid->Ny * 2 * L * nsave;
  if (sum_us_dev && nsave) {
    smm_tic(ctx, SMM_K_REDUCE);
    SMM_TRY(smm_launch_ensemble_sum(ctx, E.d_us, precision, R, per_real, sum_us_dev));
    smm_toc(ctx, SMM_K_REDUCE);
  }
  if (sum_zs_dev && nsave) SMM_TRY(smm_launch_small_sum(ctx, E.d_zs, R, (int)(nsave * M * 2), sum_zs_dev));
  if (sum_ys_dev && nsave) SMM_TRY(smm_launch_small_sum(ctx, E.d_ys, R, (int)(nsave * L * 2), sum_ys_dev));
  return SMM_OK;
}

extern "C" int smm_kde(smm_ctx* ctx, int32_t n, int32_t L, int32_t precision, const double* x, const int32_t* follow,
                       const int32_t* state, double kde_sigma, const smm_grid* grid, double* us) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_kde: ctx is NULL");
  SMM_REQUIRE(ctx, n >= 1 && L >= 1 && L <= 256 && x && follow && state && us, "smm_kde: bad argument");
  SMM_REQUIRE(ctx, precision == SMM_PREC_F32 || precision == SMM_PREC_F64, "smm_kde: unknown precision");
  SMM_REQUIRE(ctx, kde_sigma > 0, "smm_kde: kde_sigma must be > 0");
  SMM_TRY(check_grid(ctx, grid, "smm_kde"));
  for (int i = 0; i < n; i++) {
    SMM_REQUIRE(ctx, follow[i] >= 0 && follow[i] < L, "smm_kde: follow[%d]=%d outside 0..L-1", i, follow[i]);
    SMM_REQUIRE(ctx, state[i] == 1 || state[i] == -1, "smm_kde: state[%d]=%d is not +-1", i, state[i]);
  }
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t rs = precision == SMM_PREC_F64 ? 8 : 4, per = (size_t)grid->Nx * grid->Ny * 2 * L;
  void *dx, *dst32, *dfol, *dus;
  signed char* dst;
  SMM_TRY(upload_real(ctx, S_XS, x, (size_t)2 * n, precision, &dx));
  SMM_TRY(upload(ctx, S_TMP0, state, (size_t)n * 4, &dst32));
  SMM_TRY(smm_scratch(ctx, S_STATE, n, (void**)&dst));
  SMM_TRY(smm_launch_convert_i32_to_i8(ctx, (const int32_t*)dst32, dst, n));
  SMM_TRY(upload(ctx, S_FOLS, follow, (size_t)n * 4, &dfol));
  SMM_TRY(smm_scratch(ctx, S_US, per * rs, &dus));
  KdeArgs Kd;
  Kd.n = n; Kd.L = L; Kd.S = 1; Kd.Nx = grid->Nx; Kd.Ny = grid->Ny;
  Kd.gx0 = grid->x0; Kd.gy0 = grid->y0; Kd.dx = grid->dx; Kd.dy = grid->dy; Kd.sigma = kde_sigma;
  Kd.xs = dx; Kd.follows = (const int32_t*)dfol; Kd.state = dst; Kd.nsave = 1; Kd.us = dus;
  SMM_TRY(smm_launch_kde(ctx, Kd, precision));
  SMM_TRY(download_real(ctx, dus, per, precision, us));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SMM_OK;
}

// ---------------------------------------------------------------------------------------
// large resident population (C5)
// ---------------------------------------------------------------------------------------
// The persisting-L2 carve-out is a per-device limit shared by every population (and every other user of the
// device in this process): the value found before the first population raised it is restored when the last
// population that holds a share is freed.
static struct { int users; size_t saved_limit; } g_l2_carveout[64];

static void pop_free(smm_abm_pop* p) {
  if (!p) return;
  if (p->l2_persist_bytes > 0 && p->device >= 0 && p->device < 64) {
    auto& c = g_l2_carveout[p->device];
    if (--c.users == 0) {  // last user: give the carve-out back
      cudaCtxResetPersistingL2Cache();
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c.saved_limit);
    }
  }
  for (int i = 0; i < 2; i++) { if (p->x[i]) cudaFree(p->x[i]); if (p->ent[i]) cudaFree(p->ent[i]); }
  for (int i = 0; i < 3; i++) if (p->acc[i]) cudaFree(p->acc[i]);
  if (p->follow) cudaFree(p->follow);
  if (p->state) cudaFree(p->state);
  if (p->rowptr) cudaFree(p->rowptr);
  if (p->colidx) cudaFree(p->colidx);
  if (p->slice_ptr) cudaFree(p->slice_ptr);
  if (p->fallback_count) cudaFree(p->fallback_count);
  delete p;
}

#define POP_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      pop_free(p);                                                                                      \
      cudaGetLastError();                                                                               \
      return smm_fail(ctx, e__ == cudaErrorMemoryAllocation ? SMM_ERR_NOMEM : SMM_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(e__)); \
    }                                                                                                   \
  } while (0)
#define POP_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != SMM_OK) { pop_free(p); return rc__; } \
  } while (0)

extern "C" int smm_abm_pop_create(smm_ctx* ctx, const smm_params* q, int32_t precision, uint64_t seed, const double* x0,
                                  const int32_t* state, const int32_t* follow0, const double* inf0, const double* media0,
                                  const int32_t* rowptr, const int32_t* colidx, smm_abm_pop** out) {
  SMM_REQUIRE(ctx, ctx != nullptr && out != nullptr, "smm_abm_pop_create: NULL argument");
  *out = nullptr;
  SMM_TRY(check_params(ctx, q, "smm_abm_pop_create"));
  SMM_REQUIRE(ctx, precision == SMM_PREC_F32 || precision == SMM_PREC_F64, "smm_abm_pop_create: unknown precision");
  SMM_REQUIRE(ctx, x0 && state && follow0 && inf0 && media0, "smm_abm_pop_create: initial condition missing");
  SMM_REQUIRE(ctx, (rowptr == nullptr) == (colidx == nullptr), "smm_abm_pop_create: rowptr/colidx mismatch");
  const size_t n = q->n, L = q->L, M = q->M;
  for (size_t i = 0; i < n; i++) {
    SMM_REQUIRE(ctx, follow0[i] >= 0 && follow0[i] < (int)L, "smm_abm_pop_create: follow0[%zu]=%d outside 0..L-1", i, follow0[i]);
    SMM_REQUIRE(ctx, state[i] == 1 || state[i] == -1, "smm_abm_pop_create: state[%zu]=%d is not +-1", i, state[i]);
  }
  size_t nnz = 0;
  if (rowptr) {
    SMM_REQUIRE(ctx, rowptr[0] == 0, "smm_abm_pop_create: rowptr[0] must be 0");
    for (size_t j = 0; j < n; j++) SMM_REQUIRE(ctx, rowptr[j + 1] >= rowptr[j], "smm_abm_pop_create: rowptr not monotone at %zu", j);
    nnz = rowptr[n];
    for (size_t e = 0; e < nnz; e++) SMM_REQUIRE(ctx, colidx[e] >= 0 && colidx[e] < (int)n, "smm_abm_pop_create: colidx[%zu] out of range", e);
  }
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  smm_abm_pop* p = new smm_abm_pop();
  memset(p, 0, sizeof *p);
  p->device = -1;
  p->k = make_abm_const(q);
  p->precision = precision; p->seed = seed; p->step = 0; p->has_net = rowptr != nullptr; p->nnz = nnz;
  // fixed-point scale: |sum| <= n * max|x| must stay below 2^62
  double maxabs = std::max(std::max(fabs(q->dom_lo[0]), fabs(q->dom_hi[0])), std::max(fabs(q->dom_lo[1]), fabs(q->dom_hi[1])));
  int shift = 61 - (int)ceil(log2((double)n * std::max(maxabs, 1.0)));
  if (shift > 50) shift = 50;
  p->fx_scale = ldexp(1.0, shift);
  { const char* ev = getenv("SMM_POP_DECISION"); p->prefix_decision = (ev && strcmp(ev, "coop") == 0) ? 0 : 1; }
  {
    // L2 persistence window for the position gathers (SMM_POP_L2_PERSIST=0 disables): 1.46 ms per step with it,
    // 2.77 ms without at n = 1e7.  The device-wide persisting carve-out it needs shrinks the L2 left to every
    // other kernel (the PDE workloads ran 19 % slower after a population run), so it is given back when the
    // population is destroyed (pop_free).
    const char* ev = getenv("SMM_POP_L2_PERSIST");
    cudaDeviceProp prop;
    p->l2_persist_bytes = p->l2_window_max = 0;
    if (!(ev && ev[0] == '0') && rowptr && cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess &&
        prop.persistingL2CacheMaxSize > 0) {
      size_t want = std::min((size_t)prop.persistingL2CacheMaxSize, n * 2 * (precision == SMM_PREC_F64 ? (size_t)8 : (size_t)4));
      size_t cur = 0;
      if (ctx->device < 64 && cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess) {
        auto& c = g_l2_carveout[ctx->device];
        if (c.users == 0) c.saved_limit = cur;
        if (cur >= want || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
          c.users++;
          p->device = ctx->device;
          p->l2_persist_bytes = want;
          p->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        } else {
          cudaGetLastError();
        }
      } else {
        cudaGetLastError();
      }
    }
  }
  const size_t rs = precision == SMM_PREC_F64 ? 8 : 4;
  for (int i = 0; i < 2; i++) { POP_CUDA(cudaMalloc(&p->x[i], n * 2 * rs)); POP_CUDA(cudaMalloc((void**)&p->ent[i], 2 * (L + M) * 8)); }
  for (int i = 0; i < 3; i++) { POP_CUDA(cudaMalloc((void**)&p->acc[i], 2 * L * 3 * 8)); POP_CUDA(cudaMemsetAsync(p->acc[i], 0, 2 * L * 3 * 8, ctx->stream)); }
  POP_CUDA(cudaMalloc((void**)&p->follow, n));
  POP_CUDA(cudaMalloc((void**)&p->state, n));
  POP_CUDA(cudaMalloc((void**)&p->fallback_count, 16));  // + the work counter of pop_step_kernel
  POP_CUDA(cudaMemsetAsync(p->fallback_count, 0, 16, ctx->stream));
  if (rowptr) {
    // neighbour lists are kept in sliced-ELL form (slices of 32 agents, width = max degree of the slice)
    const size_t nslices = (n + 31) / 32;
    std::vector<int64_t> sp(nslices + 1, 0);
    for (size_t sl = 0; sl < nslices; sl++) {
      int w = 0;
      for (size_t j = sl * 32; j < std::min(n, sl * 32 + 32); j++) w = std::max(w, rowptr[j + 1] - rowptr[j]);
      sp[sl + 1] = sp[sl] + (int64_t)w * 32;
    }
    const size_t sell_count = std::max((size_t)sp[nslices], (size_t)1);
    POP_CUDA(cudaMalloc((void**)&p->rowptr, (n + 1) * 4));
    POP_CUDA(cudaMalloc((void**)&p->colidx, sell_count * 4));
    POP_CUDA(cudaMalloc((void**)&p->slice_ptr, (nslices + 1) * 8));
    POP_CUDA(cudaMemcpyAsync(p->rowptr, rowptr, (n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    POP_CUDA(cudaMemcpyAsync(p->slice_ptr, sp.data(), (nslices + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    void* csr;
    POP_TRY(upload(ctx, S_COLIDX, colidx, std::max(nnz, (size_t)1) * 4, &csr));
    POP_TRY(smm_launch_csr_to_sell(ctx, (int)n, p->rowptr, (const int32_t*)csr, p->slice_ptr, p->colidx));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));  // sp goes out of scope
  }
  void* tmp;
  POP_TRY(upload(ctx, S_STAGE, x0, n * 2 * 8, &tmp));
  POP_TRY(smm_launch_pack_xy(ctx, (const double*)tmp, p->x[0], n, precision));
  POP_TRY(upload(ctx, S_TMP0, state, n * 4, &tmp));
  POP_TRY(smm_launch_convert_i32_to_i8(ctx, (const int32_t*)tmp, p->state, n));
  POP_TRY(upload(ctx, S_TMP1, follow0, n * 4, &tmp));
  POP_TRY(smm_launch_convert_i32_to_u8(ctx, (const int32_t*)tmp, p->follow, n));
  std::vector<double> ent(2 * (L + M));
  for (size_t i = 0; i < L; i++) { ent[i] = inf0[i]; ent[(L + M) + i] = inf0[L + i]; }
  for (size_t i = 0; i < M; i++) { ent[L + i] = media0[i]; ent[(L + M) + L + i] = media0[M + i]; }
  POP_CUDA(cudaMemcpyAsync(p->ent[0], ent.data(), ent.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  POP_TRY(smm_launch_pop_accumulate(ctx, p, 0, 0));
  POP_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = p;
  return SMM_OK;
}

extern "C" int smm_abm_pop_step(smm_ctx* ctx, smm_abm_pop* p, int32_t nsteps) {
  SMM_REQUIRE(ctx, ctx && p, "smm_abm_pop_step: NULL argument");
  SMM_REQUIRE(ctx, nsteps >= 0, "smm_abm_pop_step: nsteps < 0");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  reset_times(ctx);
  smm_tic(ctx, SMM_K_ABM);
  for (int s = 0; s < nsteps; s++) SMM_TRY(smm_launch_pop_step(ctx, p, nullptr, nullptr, nullptr));
  smm_toc(ctx, SMM_K_ABM);
  return SMM_OK;
}

extern "C" int smm_abm_pop_step_injected(smm_ctx* ctx, smm_abm_pop* p, const double* xi, const double* u1, const double* u2) {
  SMM_REQUIRE(ctx, ctx && p && xi && u1 && u2, "smm_abm_pop_step_injected: NULL argument");
  // the entity noise of abm.jl:181-196 has no injected form here: refuse rather than silently drop it
  if (p->k.sigmahat != 0.0 || p->k.sigmatilde != 0.0)
    return smm_fail(ctx, SMM_ERR_UNSUPPORTED, "smm_abm_pop_step_injected: sigmahat=%g / sigmatilde=%g: injected influencer / media noise is "
                    "not supported for populations (use smm_abm_solve with xi_inf / xi_med, or Philox steps)", p->k.sigmahat, p->k.sigmatilde);
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = p->k.n;
  void *dxi, *du1, *du2;
  SMM_TRY(upload_real(ctx, S_XI, xi, 2 * n, p->precision, &dxi));
  SMM_TRY(upload(ctx, S_U1, u1, n * 8, &du1));
  SMM_TRY(upload(ctx, S_U2, u2, n * 8, &du2));
  SMM_TRY(smm_launch_pop_step(ctx, p, dxi, (const double*)du1, (const double*)du2));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SMM_OK;
}

extern "C" int smm_abm_pop_read(smm_ctx* ctx, smm_abm_pop* p, double* x, int32_t* follow, double* inf, double* media,
                                int64_t* numbers) {
  SMM_REQUIRE(ctx, ctx && p, "smm_abm_pop_read: NULL argument");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = p->k.n, L = p->k.L, M = p->k.M;
  const int cur = (int)(p->step & 1);
  void* tmp;
  if (x) {
    SMM_TRY(smm_scratch(ctx, S_STAGE, n * 2 * 8, &tmp));
    SMM_TRY(smm_launch_unpack_xy(ctx, p->x[cur], (double*)tmp, n, p->precision));
    SMM_TRY(download(ctx, tmp, n * 2 * 8, x));
  }
  if (follow) {
    SMM_TRY(smm_scratch(ctx, S_TMP1, n * 4, &tmp));
    SMM_TRY(smm_launch_convert_u8_to_i32(ctx, p->follow, (int32_t*)tmp, n));
    SMM_TRY(download(ctx, tmp, n * 4, follow));
  }
  std::vector<double> ent(2 * (L + M));
  std::vector<long long> acc(2 * L * 3);
  SMM_TRY(download(ctx, p->ent[cur], ent.size() * 8, ent.data()));
  SMM_TRY(download(ctx, p->acc[p->step % 3], acc.size() * 8, acc.data()));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SMM_TRY(smm_collect_times(ctx));
  if (inf) for (size_t i = 0; i < L; i++) { inf[i] = ent[i]; inf[L + i] = ent[(L + M) + i]; }
  if (media) for (size_t i = 0; i < M; i++) { media[i] = ent[L + i]; media[M + i] = ent[(L + M) + L + i]; }
  if (numbers) for (size_t j = 0; j < L; j++) for (size_t i = 0; i < M; i++) numbers[j * M + i] = acc[(j * 2 + i) * 3 + 2];
  return SMM_OK;
}

extern "C" int smm_abm_pop_stats(smm_ctx* ctx, smm_abm_pop* p, int64_t* steps, int64_t* fp64_reevaluations) {
  SMM_REQUIRE(ctx, ctx && p, "smm_abm_pop_stats: NULL argument");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long fb = 0;
  SMM_CUDA(ctx, cudaMemcpyAsync(&fb, p->fallback_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
  SMM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (steps) *steps = p->step;
  if (fp64_reevaluations) *fp64_reevaluations = (int64_t)fb;
  return SMM_OK;
}

extern "C" int smm_abm_pop_destroy(smm_ctx* ctx, smm_abm_pop* p) {
  SMM_REQUIRE(ctx, ctx != nullptr, "smm_abm_pop_destroy: ctx is NULL");
  if (p) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); pop_free(p); }
  return SMM_OK;
}
