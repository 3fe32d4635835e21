// smm_common.cuh -- shared host/device plumbing of libsmm_b200.so (sm_100a only).
#pragma once
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>

#include "../../include/smm_b200.h"

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
enum { SMM_K_ABM = 0, SMM_K_KDE = 1, SMM_K_REDUCE = 2, SMM_K_CONV = 3, SMM_K_STENCIL = 4, SMM_K_MOMENTS = 5, SMM_K_NCLASS = 8 };
enum { SMM_NSCRATCH = 48 };

struct smm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  int64_t launches = 0;
  double kernel_ms[SMM_K_NCLASS];
  cudaEvent_t ev0[SMM_K_NCLASS], ev1[SMM_K_NCLASS];
  bool ev_used[SMM_K_NCLASS];
  void* scratch[SMM_NSCRATCH];
  size_t scratch_bytes[SMM_NSCRATCH];
  void* pinned = nullptr;  // staging for pageable <-> device pipelines
  size_t pinned_bytes = 0;
  cudaStream_t copy_stream = nullptr;  // device->host copies overlapped with compute (chunked ensembles)
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  // FP32 outputs widened to Float64 by host threads (host_widen.cpp): pinned ring + its events
  struct smm_hostpool* hostpool = nullptr;
  void* h_ring = nullptr;
  size_t h_ring_bytes = 0;
  cudaEvent_t ev_h[32] = {};  // one per slot of h_ring
  int host_threads = 0;            // host widening threads (0 = default: cores / local ranks, <= 16)
  // multi-GPU (multi.cu): a leader context owns one member context per GPU; members run in shard mode
  struct smm_multi* multi = nullptr;
  bool shard_mode = false;         // member of a multi-GPU context: ensemble sums stay in device scratch for the all-reduce
};

extern std::string g_smm_last_error;

inline int smm_fail(smm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_smm_last_error = buf;
  return code;
}

#define SMM_CUDA(ctx, expr)                                                                       \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return smm_fail((ctx), e__ == cudaErrorMemoryAllocation ? SMM_ERR_NOMEM : SMM_ERR_CUDA,     \
                      "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));     \
  } while (0)

#define SMM_REQUIRE(ctx, cond, ...)                                     \
  do {                                                                  \
    if (!(cond)) return smm_fail((ctx), SMM_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

#define SMM_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != SMM_OK) return rc__; \
  } while (0)

// persistent per-context scratch buffers (grown on demand, freed with the context)
int smm_scratch(smm_ctx* ctx, int slot, size_t bytes, void** out);
int smm_pinned(smm_ctx* ctx, size_t bytes, void** out);
// bracket a kernel class with CUDA events on the context's stream
void smm_tic(smm_ctx* ctx, int cls);
void smm_toc(smm_ctx* ctx, int cls);
int smm_collect_times(smm_ctx* ctx);

// scratch slot names
enum {
  S_X0 = 0, S_STATE, S_FOLLOW, S_INF0, S_MED0, S_XI, S_U1, S_U2, S_XIINF, S_XIMED, S_SAVESTEPS,
  S_XS, S_FOLS, S_INFS, S_MEDS, S_YS, S_ZS, S_NUMBERS, S_US, S_SUMUS, S_SUMZS, S_SUMYS, S_STAGE,
  S_ROWPTR, S_COLIDX, S_TMP0, S_TMP1, S_TMP2, S_TMP3, S_TMP4, S_TMP5, S_PDE0, S_STAGE_A, S_STAGE_B, S_YS_ALL, S_ZS_ALL, S_HANDOFF
};

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
template <typename real> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; it blocks at pdl_wait() until the
// predecessor has completed and its memory is visible.  pdl_launch_dependents() lets the successor's CTAs be
// scheduled as soon as every CTA of this grid has reached it.  Both are no-ops for plainly launched kernels.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t smm_launch_pdl(int pdl_sms /* SM count of the device; 0 = plain launch */, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  // Only for grids of at most 1.5 CTAs per SM: the successor's CTAs become resident (and hold registers / shared memory at
  // pdl_wait) while this grid is still running, which pays for latency-bound launches (C3: 96.5 -> 89.2 us per
  // RK4 step) and costs throughput for large ones (177^2 x 16: 0.41 -> 0.46 ms when applied everywhere).
  const size_t ctas = (size_t)grid.x * grid.y * grid.z;
  cfg.attrs = at; cfg.numAttrs = (pdl_sms > 0 && 2 * ctas <= 3 * (size_t)pdl_sms) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 — two FP32 operations per issue slot).
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }  // scalar broadcast (folds into the operand)
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
// exp(-sqrt(r2)) -- the interaction weight of abm.jl:59 / pde.jl:17.
// FP32: two MUFU ops (sqrt.approx, ex2.approx), rel. error ~2e-7; sqrt.approx(0) = 0 so the
// self-pair gives w = 1 exactly (abm.jl:42 self-loop).  FP64: libm-accurate.
__device__ __forceinline__ float w_of_r2(float r2) {
  // (r2 * rsqrt(r2) instead of sqrt was measured 6 % slower: MUFU.SQRT is not the limiter)
  float r, w;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(r * -1.4426950408889634f));
  return w;
}
__device__ __forceinline__ double w_of_r2(double r2) { return exp(-sqrt(r2)); }

// Philox4x32-10 (Salmon et al. SC'11); oracle/philox.py restates it and is checked against
// the Random123 known-answer vectors.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

enum { SMM_STREAM_AGENT = 0, SMM_STREAM_ENTITY = 1, SMM_STREAM_INIT = 2 };

__device__ __forceinline__ float u24(uint32_t bits) { return __fmul_rn(__uint2float_rn(bits >> 8), 0x1p-24f); }

// Box-Muller on two 24-bit uniforms; explicit _rn intrinsics so that every kernel that
// inlines this produces bit-identical draws (smm_abm_draws dumps them for the oracle).
__device__ __forceinline__ void box_muller(uint32_t b0, uint32_t b1, float& n0, float& n1) {
  // hardware transcendental forms (lg2 / sin / cos on the MUFU pipe): ~100 fewer issue slots per draw
  // than logf + sincospif; their ~5e-7 absolute error is irrelevant for N(0,1) noise, and being plain
  // MUFU ops they are bit-reproducible in every kernel that inlines this function
  float ua = __fmul_rn(__fadd_rn(__uint2float_rn(b0 >> 8), 0.5f), 0x1p-24f);          // (0,1)
  float ang = __fmul_rn(__uint2float_rn(b1 >> 8), 6.283185307179586f * 0x1p-24f);     // [0, 2 pi)
  // ua < 1, but the absolute error of the hardware logarithm near 1 exceeds |log(1 - 2^-24)|: clamp so that a
  // result of +0 or above can never put a NaN into the state
  float rad = __fsqrt_rn(fmaxf(__fmul_rn(-2.0f, __logf(ua)), 0.0f));
  float s, c;
  __sincosf(ang, &s, &c);
  n0 = __fmul_rn(rad, c);
  n1 = __fmul_rn(rad, s);
}

// the four draws of one agent in one step (abm.jl:173 randn(n,2); :122 rand(); :126 rand())
__device__ __forceinline__ void agent_draws(uint64_t seed, uint32_t agent, uint32_t step, uint32_t realization,
                                            float& xi0, float& xi1, float& u1, float& u2) {
  uint4 r = philox4x32_10(make_uint4(agent, step, SMM_STREAM_AGENT, realization),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  box_muller(r.x, r.y, xi0, xi1);
  u1 = u24(r.z);
  u2 = u24(r.w);
}

// noise on influencer/medium `entity` (0..L-1 influencers, L.. media) -- abm.jl:181-196
__device__ __forceinline__ void entity_draws(uint64_t seed, uint32_t entity, uint32_t step, uint32_t realization,
                                             float& z0, float& z1) {
  uint4 r = philox4x32_10(make_uint4(entity, step, SMM_STREAM_ENTITY, realization),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  box_muller(r.x, r.y, z0, z1);
}

// single reflection (abm.jl:139-150)
__device__ __forceinline__ double reflect1(double v, double lo, double hi) {
  if (v > hi) v = -v + 2.0 * hi;
  else if (v < lo) v = -v + 2.0 * lo;
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// scalar model constants passed by value to the ABM kernels
// ---------------------------------------------------------------------------------------
struct AbmConst {
  int n, L, M;
  double eta, a, b, c, sigma, sigmahat, sigmatilde, frictionI, frictionM, dt;
  double lo0, hi0, lo1, hi1;
};

inline AbmConst make_abm_const(const smm_params* q) {
  AbmConst k;
  k.n = q->n; k.L = q->L; k.M = q->M;
  k.eta = q->eta; k.a = q->a; k.b = q->b; k.c = q->c; k.sigma = q->sigma;
  k.sigmahat = q->sigmahat; k.sigmatilde = q->sigmatilde;
  k.frictionI = q->frictionI; k.frictionM = q->frictionM; k.dt = q->dt;
  k.lo0 = q->dom_lo[0]; k.hi0 = q->dom_hi[0]; k.lo1 = q->dom_lo[1]; k.hi1 = q->dom_hi[1];
  return k;
}

#define SMM_THETA 0.1  // abm.jl:94, pde.jl:27-31
