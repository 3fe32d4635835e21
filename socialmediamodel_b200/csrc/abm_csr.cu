// abm_csr.cu -- K2/K3: one large ABM population resident in HBM (config C5: n = 1e7 agents,
// CSR follower network of out-degree ~17, L = 64 influencers).  One fused kernel per time step:
//
//   prologue (every CTA, redundantly): turn the fixed-point sums accumulated by the PREVIOUS
//       launch (sums of x_k by (influencer, medium) class) into inf_{k+1}, media_{k+1} and the
//       follower fractions (abm.jl:96-100,175-198), reflect (abm.jl:202-203);
//   main (one thread per agent): CSR gather of the neighbours' positions -> attraction
//       (abm.jl:47-67), influence (abm.jl:69-88), Philox Euler-Maruyama step (abm.jl:173),
//       reflection (abm.jl:201), influencer switch (abm.jl:103-133; FP32 with a guard band, FP64
//       re-evaluation inside the band);
//   epilogue: class sums of (x_{k+1}, follow_{k+1}) -> shared-memory bins -> 64-bit fixed-point
//       global atomics (order-independent, hence run-to-run deterministic) for the NEXT launch.
//
// No grid-wide barrier and no host round trip; three accumulator buffers rotate (read k%3,
// write (k+1)%3, clear (k+2)%3).  HBM-bound: 23 + 12*deg bytes per agent-step.
#include <string.h>

#include "abm_device.cuh"
#include "smm_internal.h"

namespace {

constexpr int kPopThreads = 256;
constexpr int kPopChunk = 8;  // 32-agent slices a warp claims at a time (one atomic per 256 agents)

struct PopArgs {
  AbmConst k;
  uint64_t seed;
  uint32_t step;
  double fx_scale;  // fixed-point scale 2^shift
  const void* x_old;
  void* x_new;
  unsigned char* follow;
  const signed char* state;
  const int32_t* rowptr;  // null = all-to-all
  const int32_t* colidx;  // SELL-32: colidx[slice_ptr[j/32] + t*32 + (j%32)] = t-th neighbour of agent j
  const int64_t* slice_ptr;
  const double* ent_old;  // [2][L+M]: x block then y block
  double* ent_new;
  const long long* acc_read;  // [2L][3]
  long long* acc_write;
  long long* acc_clear;
  const void* xi;    // injected: real [2][n] or null (Philox)
  const double* u1;  // injected f64 [n]
  const double* u2;
  unsigned long long* fallback_count;  // diagnostics: agents that took the FP64 re-evaluation
  int prefix_decision;                 // 1: prefix-sum/binary-search decision, 0: cooperative CDF walk
  unsigned int* tile_counter;          // next unclaimed chunk of kPopChunk 32-agent slices (zeroed before every launch)
};

// class index c = follow*2 + (state==+1)
template <typename real>
__global__ void __launch_bounds__(kPopThreads) pop_accumulate_kernel(int n, int L, double fx_scale, const void* xv,
                                                                     const unsigned char* follow,
                                                                     const signed char* state, long long* acc) {
  using real2 = typename vec2<real>::type;
  extern __shared__ long long bins[];  // [2L][3]
  const real2* x = static_cast<const real2*>(xv);
  for (int t = threadIdx.x; t < 2 * L * 3; t += blockDim.x) bins[t] = 0;
  __syncthreads();
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < (size_t)n; j += (size_t)gridDim.x * blockDim.x) {
    real2 p = x[j];
    int c = follow[j] * 2 + (state[j] == 1 ? 1 : 0);
    atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 0]), (unsigned long long)__double2ll_rn((double)p.x * fx_scale));
    atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 1]), (unsigned long long)__double2ll_rn((double)p.y * fx_scale));
    atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 2]), 1ull);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * L * 3; t += blockDim.x)
    if (bins[t] != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&acc[t]), (unsigned long long)bins[t]);
}

// random 8/16-byte gathers have no reuse within a CTA: read-only path, do not allocate in L1 (keeps the
// colidx rows, which every thread re-reads sector by sector, resident there)
__device__ __forceinline__ float2 ld_gather(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_gather(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <typename real>
__global__ void __launch_bounds__(kPopThreads) pop_step_kernel(PopArgs A) {
  using real2 = typename vec2<real>::type;
  const AbmConst& K = A.k;
  const int n = K.n, L = K.L, M = K.M;
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) unsigned char smem[];
  long long* bins = reinterpret_cast<long long*>(smem);             // [2L][3]
  double* eox = reinterpret_cast<double*>(bins + 2 * L * 3);        // old entities [L+M]
  double* eoy = eox + (L + M);
  double* enx = eoy + (L + M);                                      // new entities [L+M]
  double* eny = enx + (L + M);
  double* frac = eny + (L + M);                                     // [L]
  float* fnx = reinterpret_cast<float*>(frac + L);                  // FP32 copies for the guarded decision
  float* fny = fnx + L;
  float* ffrac = fny + L;
  double* att_scr = reinterpret_cast<double*>(ffrac + L + (L & 1)) + (tid >> 5) * L;  // [nwarps][L]
  // (prefix variant) 16-byte aligned by construction: one pad word when the count of preceding 4-byte words is odd/8-byte misaligned
  float4* tab = reinterpret_cast<float4*>(reinterpret_cast<double*>(ffrac + L + (L & 1)) + (kPopThreads / 32) * L + ((19 * L + 4 * M + (3 * L + (L & 1)) / 2) & 1));
  float* pref = reinterpret_cast<float*>(tab + L);  // [(L+1)/2][kPopThreads]

  // ---- prologue -------------------------------------------------------------------------
  for (int t = tid; t < 2 * L * 3; t += kPopThreads) bins[t] = 0;
  if (blockIdx.x == 0)
    for (int t = tid; t < 2 * L * 3; t += kPopThreads) A.acc_clear[t] = 0;
  const double sdt = sqrt(K.dt);
  for (int t = tid; t < L + M; t += kPopThreads) {
    const bool is_inf = t < L;
    double sx, sy, cnt, ss = 0.0;
    if (is_inf) {
      long long c0 = A.acc_read[(t * 2 + 0) * 3 + 2], c1 = A.acc_read[(t * 2 + 1) * 3 + 2];
      sx = (double)(A.acc_read[(t * 2 + 0) * 3 + 0] + A.acc_read[(t * 2 + 1) * 3 + 0]) / A.fx_scale;
      sy = (double)(A.acc_read[(t * 2 + 0) * 3 + 1] + A.acc_read[(t * 2 + 1) * 3 + 1]) / A.fx_scale;
      cnt = (double)(c0 + c1);
      ss = (double)(c1 - c0);
    } else {
      const int s = t - L;  // medium row 0 <-> state -1, row 1 <-> +1 (abm.jl:189-191)
      long long ax = 0, ay = 0, c = 0;
      for (int i = 0; i < L; i++) {
        ax += A.acc_read[(i * 2 + s) * 3 + 0];
        ay += A.acc_read[(i * 2 + s) * 3 + 1];
        c += A.acc_read[(i * 2 + s) * 3 + 2];
      }
      sx = (double)ax / A.fx_scale; sy = (double)ay / A.fx_scale; cnt = (double)c;
    }
    double px = A.ent_old[t], py = A.ent_old[(L + M) + t];
    eox[t] = px; eoy[t] = py;
    const double fr = is_inf ? K.frictionI : K.frictionM;
    const double sg = is_inf ? K.sigmahat : K.sigmatilde;
    double z0 = 0.0, z1 = 0.0;
    if (sg != 0.0 && A.xi == nullptr) {
      float f0, f1;
      entity_draws(A.seed, (uint32_t)t, A.step, 0u, f0, f1);
      z0 = f0; z1 = f1;
    }
    const double nz = 1.0 / fr * sdt * sg;
    if (cnt > 0) {
      px = px + K.dt / fr * (sx / cnt - px) + nz * z0;
      py = py + K.dt / fr * (sy / cnt - py) + nz * z1;
    } else {
      px = px + nz * z0;
      py = py + nz * z1;
    }
    px = reflect1(px, K.lo0, K.hi0);
    py = reflect1(py, K.lo1, K.hi1);
    enx[t] = px; eny[t] = py;
    if (is_inf) {
      double f = ss / cnt;  // 0/0 = NaN (Q5)
      frac[t] = f;
      fnx[t] = (float)px; fny[t] = (float)py; ffrac[t] = (float)f;
      if (A.prefix_decision) {  // eta * max(+-frac, 0.1) per state (abm.jl:115-119); NaN stays NaN (Q5)
        const float ff = (float)f, et = (float)K.eta;
        tab[t] = make_float4((float)px, (float)py, ff != ff ? ff : et * fmaxf(ff, 0.1f), ff != ff ? ff : et * fmaxf(-ff, 0.1f));
      }
    }
    if (blockIdx.x == 0) { A.ent_new[t] = px; A.ent_new[(L + M) + t] = py; }
  }
  __syncthreads();

  // ---- main: one thread per agent ---------------------------------------------------------
  const real2* xo = static_cast<const real2*>(A.x_old);
  real2* xw = static_cast<real2*>(A.x_new);
  const double noise_x = sdt * K.sigma;
  unsigned int nfallback = 0;
  // Work is claimed, not assigned: a warp takes the next chunk of kPopChunk slices (32 consecutive agents each) from a global
  // counter.  With a fixed share per CTA (grid-stride loop) the CTAs of an SM finish far apart -- the warp schedulers favour
  // the older ones, see the dense kernel -- and the launch ends in a tail at falling occupancy; claimed work keeps every
  // warp busy until the population is used up.  Results do not depend on who processes a slice (per-agent arithmetic,
  // fixed-point class sums).
  const int lane = tid & 31;
  for (;;) {
  unsigned int chunk = 0;
  if (lane == 0) chunk = atomicAdd(A.tile_counter, 1u);
  chunk = __shfl_sync(0xffffffffu, chunk, 0);
  const size_t cbase = (size_t)chunk * (32 * kPopChunk);
  if (cbase >= (size_t)n) break;
  for (int sl = 0; sl < kPopChunk; sl++) {
    const size_t base = cbase + (size_t)sl * 32;        // this warp's slice
    if (base >= (size_t)n) break;
    const bool valid = base + lane < (size_t)n;         // warp-uniform loop; tail lanes idle but shuffle
    const size_t j = valid ? base + lane : (size_t)n - 1;
    const real2 xj = xo[j];
    real ax = 0, ay = 0, ws = 0;
    if (A.rowptr) {
      // sliced-ELL neighbour lists: lane l of a warp reads colidx[sp + t*32 + l] -> one coalesced
      // wavefront per neighbour slot instead of 32 (per-thread CSR rows made this kernel L1-bound)
      const int e0 = A.rowptr[j], e1 = valid ? A.rowptr[j + 1] : e0;
      const int32_t* nb = A.colidx + A.slice_ptr[base >> 5] + lane;
#pragma unroll 4
      for (int e = e0; e < e1; e++, nb += 32) {  // abm.jl:51,57-62
        real2 p = ld_gather(xo + *nb);
        real dx = p.x - xj.x, dy = p.y - xj.y;
        real w = w_of_r2(dx * dx + dy * dy);
        ws += w; ax += w * dx; ay += w * dy;
      }
      if (e1 == e0) ws = (real)1;  // abm.jl:52-53
    } else {
#pragma unroll 4
      for (int i = 0; i < n; i++) {
        real2 p = xo[i];
        real dx = p.x - xj.x, dy = p.y - xj.y;
        real w = w_of_r2(dx * dx + dy * dy);
        ws += w; ax += w * dx; ay += w * dy;
      }
    }
    const int sj = A.state[j], fj = A.follow[j];
    const double x = (double)xj.x, y = (double)xj.y;
    const int mrow = L + (sj == 1 ? 1 : 0);
    double fx = K.a * (double)(ax / ws) + K.c * (eox[mrow] - x) + K.b * (eox[fj] - x);
    double fy = K.a * (double)(ay / ws) + K.c * (eoy[mrow] - y) + K.b * (eoy[fj] - y);
    double xi0, xi1, u1, u2;
    if (A.xi) {
      const real* xp = static_cast<const real*>(A.xi);
      xi0 = (double)xp[j]; xi1 = (double)xp[(size_t)n + j];
      u1 = A.u1[j]; u2 = A.u2[j];
    } else {
      float f0, f1, g1, g2;
      agent_draws(A.seed, (uint32_t)j, A.step, 0u, f0, f1, g1, g2);
      xi0 = f0; xi1 = f1; u1 = g1; u2 = g2;
    }
    double xn0 = reflect1(x + K.dt * fx + noise_x * xi0, K.lo0, K.hi0);
    double xn1 = reflect1(y + K.dt * fy + noise_x * xi1, K.lo1, K.hi1);
    int fnew;
    if (sizeof(real) == 4 && A.xi == nullptr) {
      // Philox uniforms are 24-bit, exactly representable in FP32
      if (A.prefix_decision)
        fnew = warp_switch_decision_prefix(valid, (float)xj.x, (float)xj.y, sj, fj, (float)u1, (float)u2, tab, enx, eny, frac, L,
                                           (float)K.dt, K.eta, K.dt, 2e-5f, pref, att_scr, &nfallback);
      else
        fnew = warp_switch_decision(valid, (float)xj.x, (float)xj.y, sj, fj, (float)u1, (float)u2, fnx, fny, ffrac, enx, eny,
                                  frac, L, (float)K.eta, (float)K.dt, K.eta, K.dt, 2e-5f, att_scr, &nfallback);
    } else {
      fnew = switch_decision_f64(x, y, sj, fj, enx, eny, frac, L, K.eta, K.dt, u1, u2);
    }
    if (valid) {
      const real2 pn = mk2((real)xn0, (real)xn1);
      xw[j] = pn;
      A.follow[j] = (unsigned char)fnew;
      // class sums of the NEW state for the next launch (fixed point: order-independent)
      const int c = fnew * 2 + (sj == 1 ? 1 : 0);
      atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 0]), (unsigned long long)__double2ll_rn((double)pn.x * A.fx_scale));
      atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 1]), (unsigned long long)__double2ll_rn((double)pn.y * A.fx_scale));
      atomicAdd(reinterpret_cast<unsigned long long*>(&bins[c * 3 + 2]), 1ull);
    }
  }
  }
  __syncthreads();
  for (int t = tid; t < 2 * L * 3; t += kPopThreads)
    if (bins[t] != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&A.acc_write[t]), (unsigned long long)bins[t]);
  if (A.fallback_count && sizeof(real) == 4 && A.xi == nullptr) {
    nfallback = warp_sum((int)nfallback);
    if ((tid & 31) == 0 && nfallback) atomicAdd(A.fallback_count, (unsigned long long)nfallback);
  }
}

template <typename real>
__global__ void pack_xy_kernel(const double* __restrict__ xy, typename vec2<real>::type* __restrict__ out, size_t n) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    out[j] = mk2((real)xy[j], (real)xy[n + j]);
}
template <typename real>
__global__ void unpack_xy_kernel(const typename vec2<real>::type* __restrict__ in, double* __restrict__ xy, size_t n) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
    auto p = in[j];
    xy[j] = (double)p.x;
    xy[n + j] = (double)p.y;
  }
}
template <typename Tin, typename Tout>
__global__ void cvt_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, size_t n) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) out[j] = (Tout)in[j];
}

// CSR -> SELL-32 (one warp per slice of 32 consecutive agents; padding = the agent itself, never read)
__global__ void csr_to_sell_kernel(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr,
                                   const int64_t* __restrict__ slice_ptr, int32_t* __restrict__ sell) {
  const int lane = threadIdx.x & 31;
  const size_t slice = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t j = slice * 32 + lane;
  if (slice * 32 >= (size_t)n) return;
  const int64_t sp = slice_ptr[slice];
  const int width = (int)((slice_ptr[slice + 1] - sp) >> 5);
  const int e0 = j < (size_t)n ? rowptr[j] : 0, deg = j < (size_t)n ? rowptr[j + 1] - e0 : 0;
  for (int t = 0; t < width; t++) sell[sp + (int64_t)t * 32 + lane] = t < deg ? csr[e0 + t] : (int32_t)(j < (size_t)n ? j : 0);
}

size_t pop_smem_bytes(int L, int M, bool prefix) {
  size_t b = (size_t)2 * L * 3 * 8 + sizeof(double) * (4 * (L + M) + L) + sizeof(float) * (3 * L + 2) + sizeof(double) * (kPopThreads / 32) * L + 16;
  if (prefix) b += 16 + sizeof(float4) * L + sizeof(float) * (size_t)((L + 1) / 2) * kPopThreads;
  return b;
}

}  // namespace

static int pop_grid(smm_ctx* ctx, int n) {
  long long want = ((long long)n + kPopThreads - 1) / kPopThreads;
  long long cap = (long long)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
  return (int)(want < cap ? (want ? want : 1) : cap);
}

int smm_launch_pop_accumulate(smm_ctx* ctx, smm_abm_pop* p, int buf, int accslot) {
  const int L = p->k.L;
  double scale = p->fx_scale;
  size_t sm = (size_t)2 * L * 3 * 8;
  int grid = pop_grid(ctx, p->k.n);
  if (p->precision == SMM_PREC_F64)
    pop_accumulate_kernel<double><<<grid, kPopThreads, sm, ctx->stream>>>(p->k.n, L, scale, p->x[buf], p->follow, p->state, p->acc[accslot]);
  else
    pop_accumulate_kernel<float><<<grid, kPopThreads, sm, ctx->stream>>>(p->k.n, L, scale, p->x[buf], p->follow, p->state, p->acc[accslot]);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_pop_step(smm_ctx* ctx, smm_abm_pop* p, const void* xi, const double* u1, const double* u2) {
  PopArgs A;
  A.k = p->k;
  A.seed = p->seed;
  A.step = (uint32_t)p->step;
  A.fx_scale = p->fx_scale;
  const int cur = (int)(p->step & 1);
  A.x_old = p->x[cur]; A.x_new = p->x[cur ^ 1];
  A.follow = p->follow; A.state = p->state;
  A.rowptr = p->has_net ? p->rowptr : nullptr; A.colidx = p->has_net ? p->colidx : nullptr;
  A.slice_ptr = p->has_net ? p->slice_ptr : nullptr;
  A.ent_old = p->ent[cur]; A.ent_new = p->ent[cur ^ 1];
  A.acc_read = p->acc[p->step % 3]; A.acc_write = p->acc[(p->step + 1) % 3]; A.acc_clear = p->acc[(p->step + 2) % 3];
  A.xi = xi; A.u1 = u1; A.u2 = u2;
  A.fallback_count = p->fallback_count;
  A.prefix_decision = p->prefix_decision && p->precision != SMM_PREC_F64;
  A.tile_counter = reinterpret_cast<unsigned int*>(p->fallback_count + 1);
  SMM_CUDA(ctx, cudaMemsetAsync(A.tile_counter, 0, 4, ctx->stream));
  size_t sm = pop_smem_bytes(p->k.L, p->k.M, A.prefix_decision != 0);
  if (sm > 48 * 1024) {
    if (p->precision == SMM_PREC_F64) SMM_CUDA(ctx, cudaFuncSetAttribute(pop_step_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    else SMM_CUDA(ctx, cudaFuncSetAttribute(pop_step_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  }
  int grid = pop_grid(ctx, p->k.n);
  if (p->l2_persist_bytes > 0) {
    // keep the gathered (old) position array resident in L2; the colidx stream and the new
    // positions pass through as streaming traffic
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    size_t win = (size_t)p->k.n * 2 * (p->precision == SMM_PREC_F64 ? 8 : 4);
    if (win > p->l2_window_max) win = p->l2_window_max;
    av.accessPolicyWindow.base_ptr = const_cast<void*>(A.x_old);
    av.accessPolicyWindow.num_bytes = win;
    av.accessPolicyWindow.hitRatio = (float)(p->l2_persist_bytes >= win ? 1.0 : (double)p->l2_persist_bytes / (double)win);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  if (p->precision == SMM_PREC_F64) pop_step_kernel<double><<<grid, kPopThreads, sm, ctx->stream>>>(A);
  else pop_step_kernel<float><<<grid, kPopThreads, sm, ctx->stream>>>(A);
  if (p->l2_persist_bytes > 0) {
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  p->step++;
  return SMM_OK;
}

static unsigned strided_grid(smm_ctx* ctx, size_t n) {
  size_t g = (n + 255) / 256, cap = (size_t)ctx->sm_count * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}
int smm_launch_pack_xy(smm_ctx* ctx, const double* xy, void* x2, size_t n, int precision) {
  if (precision == SMM_PREC_F64) pack_xy_kernel<double><<<strided_grid(ctx, n), 256, 0, ctx->stream>>>(xy, (double2*)x2, n);
  else pack_xy_kernel<float><<<strided_grid(ctx, n), 256, 0, ctx->stream>>>(xy, (float2*)x2, n);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
int smm_launch_unpack_xy(smm_ctx* ctx, const void* x2, double* xy, size_t n, int precision) {
  if (precision == SMM_PREC_F64) unpack_xy_kernel<double><<<strided_grid(ctx, n), 256, 0, ctx->stream>>>((const double2*)x2, xy, n);
  else unpack_xy_kernel<float><<<strided_grid(ctx, n), 256, 0, ctx->stream>>>((const float2*)x2, xy, n);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
int smm_launch_convert_i32_to_u8(smm_ctx* ctx, const int32_t* in, unsigned char* out, size_t count) {
  cvt_kernel<int32_t, unsigned char><<<strided_grid(ctx, count), 256, 0, ctx->stream>>>(in, out, count);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
int smm_launch_convert_u8_to_i32(smm_ctx* ctx, const unsigned char* in, int32_t* out, size_t count) {
  cvt_kernel<unsigned char, int32_t><<<strided_grid(ctx, count), 256, 0, ctx->stream>>>(in, out, count);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_csr_to_sell(smm_ctx* ctx, int n, const int32_t* rowptr, const int32_t* csr, const int64_t* slice_ptr, int32_t* sell) {
  const size_t nslices = ((size_t)n + 31) / 32;
  csr_to_sell_kernel<<<(unsigned)((nslices * 32 + 255) / 256), 256, 0, ctx->stream>>>(n, rowptr, csr, slice_ptr, sell);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
