// kde.cu -- K4: kernel-density estimate of ABM snapshots on the ensemble grid, and the ensemble
// sums.  Replaces ensemble.jl:27-41 (+ sumgaussian :2-8, gaussian pde.jl:122) and the
// numerator of the ensemble means, plotting.jl:361-363.
//
// us[ix,iy,i,j] = (1/n) sum_{agents a with medium i, influencer j} G(gx_ix - x_a) G(gy_iy - y_a) / (2 pi sigma^2)
// The Gaussian is separable, so per class (i,j) the density is a rank-|class| sum of outer
// products ex[a][ix] * ey[a][iy]: (Nx+Ny) exps and Nx*Ny FMAs per agent instead of Nx*Ny exps.
// One CTA per snapshot; agents are counting-sorted by class in shared memory; each thread owns
// an 8 x 4 register tile of the grid.
#include "smm_internal.h"

namespace {

__device__ __forceinline__ void ld4(float* d, const float* s) {
  float4 a = *reinterpret_cast<const float4*>(s);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
}
__device__ __forceinline__ void ld4(double* d, const double* s) {
  double2 a = reinterpret_cast<const double2*>(s)[0], b = reinterpret_cast<const double2*>(s)[1];
  d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
}

constexpr int KC = 32;            // agents per staged chunk
constexpr int KW = 8;             // warps per CTA
constexpr int TXL = 3;            // x columns per lane (stride 32)  -> 96 columns per pass
constexpr int TYW = 12;           // y rows per warp                 -> 96 rows per pass
constexpr int SPAN = 96;

__device__ __forceinline__ float gauss1d(float d, float inv2s2) { return __expf(-d * d * inv2s2); }
__device__ __forceinline__ double gauss1d(double d, double inv2s2) { return exp(-d * d * inv2s2); }

// One CTA (8 warps) per snapshot.  Lane l of every warp owns grid columns l, l+32, l+64 of the current
// 96-column pass, warp w owns rows 12w..12w+11: per staged agent a thread reads 3 ex values
// (consecutive lanes -> conflict-free) and 12 ey values (same address in all lanes -> broadcast, as
// three 128-bit loads) for 36 FMAs, and the stores of a warp are 128-byte contiguous rows.
template <typename real>
__global__ void __launch_bounds__(KW * 32, sizeof(real) == 4 ? 4 : 1) kde_kernel(KdeArgs A) {
  const int n = A.n, L = A.L, Nx = A.Nx, Ny = A.Ny;
  const int ncls = 2 * L;
  const int s = blockIdx.x;
  const int r = s / A.nsave;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = KW * 32;

  extern __shared__ __align__(16) unsigned char smem[];
  real* ex = reinterpret_cast<real*>(smem);        // [KC][SPAN]
  real* ey = ex + (size_t)KC * SPAN;               // [KC][SPAN]
  real* ax = ey + (size_t)KC * SPAN;               // [n] agent x sorted by class
  real* ay = ax + n;                               // [n]
  int* off = reinterpret_cast<int*>(ay + n);       // [ncls+1]

  const real* xs = static_cast<const real*>(A.xs) + (size_t)s * 2 * n;
  const int32_t* fol = A.follows + (size_t)s * n;
  const signed char* st = A.state + (size_t)r * n;

  // ---- counting sort of the agents by class c = j*2 + i (i = 1 for state +1: ensemble.jl:26,36)
  for (int c = tid; c <= ncls; c += nthreads) off[c] = 0;
  __syncthreads();
  for (int a = tid; a < n; a += nthreads) atomicAdd(&off[fol[a] * 2 + (st[a] == 1 ? 1 : 0) + 1], 1);
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < ncls; c++) off[c + 1] += off[c];
  }
  __syncthreads();
  // stable placement, one warp per class: the agents are scanned in order, 32 at a time; a ballot gives every
  // member its rank within the class.  Summation order, and therefore the FP32 result, is reproducible run to run
  for (int c = warp; c < ncls; c += KW) {
    int pos = off[c];
    for (int a0 = 0; a0 < n; a0 += 32) {
      const int a = a0 + lane;
      const bool in = a < n && fol[a] * 2 + (st[a] == 1 ? 1 : 0) == c;
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) {
        const int q = pos + __popc(m & ((1u << lane) - 1u));
        ax[q] = xs[a]; ay[q] = xs[n + a];
      }
      pos += __popc(m);
    }
  }
  __syncthreads();

  const real inv2s2 = (real)(1.0 / (2.0 * A.sigma * A.sigma));
  const real scale = (real)(1.0 / (double)n / (2.0 * 3.14159265358979323846 * A.sigma * A.sigma));
  real* us = static_cast<real*>(A.us) + (size_t)s * ncls * Nx * Ny;

  for (int sy0 = 0; sy0 < Ny; sy0 += SPAN) {
    for (int sx0 = 0; sx0 < Nx; sx0 += SPAN) {
      const int iy0 = sy0 + warp * TYW;
      // staging role of this thread in this pass (same expression as before: the coordinate in Float64, then rounded)
      const bool stage_x = tid < SPAN;
      const int stage_i = stage_x ? sx0 + tid : sy0 + tid - SPAN;
      const bool stage_in = stage_x ? stage_i < Nx : stage_i < Ny;
      const real stage_coord = stage_x ? (real)(A.gx0 + stage_i * A.dx) : (real)(A.gy0 + stage_i * A.dy);
      for (int c = 0; c < ncls; c++) {
        real acc[TYW][TXL];
        float2 acc2[TYW / 2][TXL];  // FP32: rows (2b, 2b+1) as packed pairs -> FFMA2, two FMAs per issue slot
#pragma unroll
        for (int b = 0; b < TYW; b++)
#pragma unroll
          for (int a = 0; a < TXL; a++) { acc[b][a] = 0; acc2[b / 2][a] = make_float2(0.f, 0.f); }
        for (int k0 = off[c]; k0 < off[c + 1]; k0 += KC) {
          const int kc = min(KC, off[c + 1] - k0);
          __syncthreads();  // previous chunk fully consumed
          // separable factors of the chunk: thread t < 96 owns grid column t (ex), thread 96 + t grid row t (ey); its
          // coordinate sits in a register, the agent coordinate is a broadcast load: 6 instructions per factor
          if (tid < 2 * SPAN) {
            real* dst = stage_x ? ex + tid : ey + (tid - SPAN);
            const real* src = (stage_x ? ax : ay) + k0;
            for (int k = 0; k < kc; k++) dst[k * SPAN] = stage_in ? gauss1d(stage_coord - src[k], inv2s2) : (real)0;
          }
          __syncthreads();
          for (int k = 0; k < kc; k++) {
            real vx[TXL], vy[TYW];
#pragma unroll
            for (int a = 0; a < TXL; a++) vx[a] = ex[k * SPAN + lane + 32 * a];
#pragma unroll
            for (int b = 0; b < TYW; b += 4) ld4(vy + b, ey + k * SPAN + warp * TYW + b);
            if constexpr (sizeof(real) == 4) {
#pragma unroll
              for (int b = 0; b < TYW; b += 2)
#pragma unroll
                for (int a = 0; a < TXL; a++)
                  acc2[b / 2][a] = fma2(make_float2((float)vx[a], (float)vx[a]), make_float2((float)vy[b], (float)vy[b + 1]), acc2[b / 2][a]);
            } else {
#pragma unroll
              for (int b = 0; b < TYW; b++)
#pragma unroll
                for (int a = 0; a < TXL; a++) acc[b][a] += vx[a] * vy[b];
            }
          }
        }
        if constexpr (sizeof(real) == 4) {
#pragma unroll
          for (int b = 0; b < TYW; b += 2)
#pragma unroll
            for (int a = 0; a < TXL; a++) { acc[b][a] = acc2[b / 2][a].x; acc[b + 1][a] = acc2[b / 2][a].y; }
        }
        real* o = us + (size_t)c * Nx * Ny + (size_t)iy0 * Nx + sx0 + lane;
        const int nrows = min(TYW, Ny - iy0);  // warp-uniform; column predicates once per tile, not per element
        bool colin[TXL];
#pragma unroll
        for (int a = 0; a < TXL; a++) colin[a] = sx0 + lane + 32 * a < Nx;
#pragma unroll
        for (int b = 0; b < TYW; b++) {
          if (b < nrows) {
#pragma unroll
            for (int a = 0; a < TXL; a++)
              if (colin[a]) o[32 * a] = acc[b][a] * scale;
          }
          o += Nx;
        }
      }
    }
  }
}

// sum over realizations of per-realization blocks (deterministic: fixed order, f64 accumulate)
template <typename real>
__global__ void ensemble_sum_kernel(const real* __restrict__ us, int R, size_t per_real, double* __restrict__ out, int accumulate) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per_real) return;
  double acc = accumulate ? out[e] : 0.0;  // chunked ensembles add their chunk to the running sum
  for (int r = 0; r < R; r++) acc += (double)us[(size_t)r * per_real + e];
  out[e] = acc;
}

template <typename Tin, typename Tout>
__global__ void convert_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, size_t count) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) out[e] = (Tout)in[e];
}

}  // namespace

int smm_launch_kde(smm_ctx* ctx, const KdeArgs& A, int precision) {
  if (A.S <= 0) return SMM_OK;
  size_t rs = precision == SMM_PREC_F64 ? 8 : 4;
  size_t sm = rs * ((size_t)2 * KC * SPAN + 2 * (size_t)A.n) + sizeof(int) * (2 * A.L + 2) + 16;
  SMM_REQUIRE(ctx, sm <= 227 * 1024, "KDE kernel: n=%d needs %zu B of shared memory", A.n, sm);
  if (precision == SMM_PREC_F64) {
    if (sm > 48 * 1024) SMM_CUDA(ctx, cudaFuncSetAttribute(kde_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kde_kernel<double><<<A.S, KW * 32, sm, ctx->stream>>>(A);
  } else {
    if (sm > 48 * 1024) SMM_CUDA(ctx, cudaFuncSetAttribute(kde_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kde_kernel<float><<<A.S, KW * 32, sm, ctx->stream>>>(A);
  }
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_ensemble_sum(smm_ctx* ctx, const void* us, int precision, int R, size_t per_real, double* sum_us, int accumulate) {
  unsigned grid = (unsigned)((per_real + 255) / 256);
  if (precision == SMM_PREC_F64) ensemble_sum_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)us, R, per_real, sum_us, accumulate);
  else ensemble_sum_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)us, R, per_real, sum_us, accumulate);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

// few outputs, many realizations (ys / zs sums): one CTA per output element, fixed-order tree reduction
__global__ void small_sum_kernel(const double* __restrict__ v, int R, int per_real, double* __restrict__ out) {
  __shared__ double sh[256];
  const int e = blockIdx.x;
  double acc = 0.0;
  for (int r = threadIdx.x; r < R; r += 256) acc += v[(size_t)r * per_real + e];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[e] = sh[0];
}

int smm_launch_small_sum(smm_ctx* ctx, const double* v, int R, int per_real, double* out) {
  small_sum_kernel<<<per_real, 256, 0, ctx->stream>>>(v, R, per_real, out);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

static unsigned conv_grid(smm_ctx* ctx, size_t count) {
  size_t g = (count + 255) / 256;
  size_t cap = (size_t)ctx->sm_count * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

int smm_launch_convert_f64_to_real(smm_ctx* ctx, const double* in, void* out, size_t count, int precision) {
  if (!count) return SMM_OK;
  if (precision == SMM_PREC_F64) convert_kernel<double, double><<<conv_grid(ctx, count), 256, 0, ctx->stream>>>(in, (double*)out, count);
  else convert_kernel<double, float><<<conv_grid(ctx, count), 256, 0, ctx->stream>>>(in, (float*)out, count);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_convert_real_to_f64(smm_ctx* ctx, const void* in, double* out, size_t count, int precision) {
  if (!count) return SMM_OK;
  if (precision == SMM_PREC_F64) convert_kernel<double, double><<<conv_grid(ctx, count), 256, 0, ctx->stream>>>((const double*)in, out, count);
  else convert_kernel<float, double><<<conv_grid(ctx, count), 256, 0, ctx->stream>>>((const float*)in, out, count);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_convert_i32_to_i8(smm_ctx* ctx, const int32_t* in, signed char* out, size_t count) {
  if (!count) return SMM_OK;
  convert_kernel<int32_t, signed char><<<conv_grid(ctx, count), 256, 0, ctx->stream>>>(in, out, count);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
