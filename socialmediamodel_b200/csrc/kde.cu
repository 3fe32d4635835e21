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

constexpr int KC = 32;  // agents per staged chunk
constexpr int TX = 8, TY = 4;

__device__ __forceinline__ float gauss1d(float d, float inv2s2) { return __expf(-d * d * inv2s2); }
__device__ __forceinline__ double gauss1d(double d, double inv2s2) { return exp(-d * d * inv2s2); }

template <typename real>
__global__ void kde_kernel(KdeArgs A) {
  const int n = A.n, L = A.L, Nx = A.Nx, Ny = A.Ny;
  const int ncls = 2 * L;
  const int s = blockIdx.x;
  const int r = s / A.nsave;
  const int txn = blockDim.x, tyn = blockDim.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * txn + tx, nthreads = txn * tyn;
  const int NxP = (Nx + TX - 1) / TX * TX, NyP = (Ny + TY - 1) / TY * TY;

  extern __shared__ __align__(16) unsigned char smem[];
  real* ex = reinterpret_cast<real*>(smem);        // [KC][NxP]
  real* ey = ex + (size_t)KC * NxP;                // [KC][NyP]
  real* ax = ey + (size_t)KC * NyP;                // [n] agent x sorted by class
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
  // stable placement (one thread per class scans the agents in order): summation order, and
  // therefore the FP32 result, is reproducible run to run
  for (int c = tid; c < ncls; c += nthreads) {
    int pos = off[c];
    for (int a = 0; a < n; a++) {
      if (fol[a] * 2 + (st[a] == 1 ? 1 : 0) == c) { ax[pos] = xs[a]; ay[pos] = xs[n + a]; pos++; }
    }
  }
  __syncthreads();

  const real inv2s2 = (real)(1.0 / (2.0 * A.sigma * A.sigma));
  const real scale = (real)(1.0 / (double)n / (2.0 * 3.14159265358979323846 * A.sigma * A.sigma));
  real* us = static_cast<real*>(A.us) + (size_t)s * ncls * Nx * Ny;

  for (int sy0 = 0; sy0 < NyP; sy0 += tyn * TY) {
    for (int sx0 = 0; sx0 < NxP; sx0 += txn * TX) {
      const int ix0 = sx0 + tx * TX, iy0 = sy0 + ty * TY;
      const bool active = ix0 < NxP && iy0 < NyP;
      for (int c = 0; c < ncls; c++) {
        real acc[TY][TX];
#pragma unroll
        for (int b = 0; b < TY; b++)
#pragma unroll
          for (int a = 0; a < TX; a++) acc[b][a] = 0;
        for (int k0 = off[c]; k0 < off[c + 1]; k0 += KC) {
          const int kc = min(KC, off[c + 1] - k0);
          __syncthreads();  // previous chunk fully consumed
          for (int e = tid; e < kc * NxP; e += nthreads) {
            int k = e / NxP, ix = e - k * NxP;
            ex[e] = ix < Nx ? gauss1d((real)(A.gx0 + ix * A.dx) - ax[k0 + k], inv2s2) : (real)0;
          }
          for (int e = tid; e < kc * NyP; e += nthreads) {
            int k = e / NyP, iy = e - k * NyP;
            ey[e] = iy < Ny ? gauss1d((real)(A.gy0 + iy * A.dy) - ay[k0 + k], inv2s2) : (real)0;
          }
          __syncthreads();
          if (active) {
            for (int k = 0; k < kc; k++) {
              real vx[TX], vy[TY];
#pragma unroll
              for (int a = 0; a < TX; a++) vx[a] = ex[k * NxP + ix0 + a];
#pragma unroll
              for (int b = 0; b < TY; b++) vy[b] = ey[k * NyP + iy0 + b];
#pragma unroll
              for (int b = 0; b < TY; b++)
#pragma unroll
                for (int a = 0; a < TX; a++) acc[b][a] += vx[a] * vy[b];
            }
          }
        }
        if (active) {
          real* o = us + (size_t)c * Nx * Ny;
#pragma unroll
          for (int b = 0; b < TY; b++)
#pragma unroll
            for (int a = 0; a < TX; a++)
              if (ix0 + a < Nx && iy0 + b < Ny) o[(size_t)(iy0 + b) * Nx + ix0 + a] = acc[b][a] * scale;
        }
      }
    }
  }
}

// sum over realizations of per-realization blocks (deterministic: fixed order, f64 accumulate)
template <typename real>
__global__ void ensemble_sum_kernel(const real* __restrict__ us, int R, size_t per_real, double* __restrict__ out) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per_real) return;
  double acc = 0.0;
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
  int txn = (A.Nx + TX - 1) / TX;
  if (txn > 32) txn = 32;
  int tyn = (A.Ny + TY - 1) / TY;
  int maxy = 512 / txn;
  if (tyn > maxy) tyn = maxy;
  size_t rs = precision == SMM_PREC_F64 ? 8 : 4;
  int NxP = (A.Nx + TX - 1) / TX * TX, NyP = (A.Ny + TY - 1) / TY * TY;
  size_t sm = rs * ((size_t)KC * (NxP + NyP) + 2 * (size_t)A.n) + sizeof(int) * (4 * A.L + 2) + 16;
  SMM_REQUIRE(ctx, sm <= 227 * 1024, "KDE kernel: n=%d / grid %dx%d need %zu B of shared memory", A.n, A.Nx, A.Ny, sm);
  dim3 block(txn, tyn);
  if (precision == SMM_PREC_F64) {
    SMM_CUDA(ctx, cudaFuncSetAttribute(kde_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kde_kernel<double><<<A.S, block, sm, ctx->stream>>>(A);
  } else {
    SMM_CUDA(ctx, cudaFuncSetAttribute(kde_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kde_kernel<float><<<A.S, block, sm, ctx->stream>>>(A);
  }
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_ensemble_sum(smm_ctx* ctx, const void* us, int precision, int R, size_t per_real, double* sum_us) {
  unsigned grid = (unsigned)((per_real + 255) / 256);
  if (precision == SMM_PREC_F64) ensemble_sum_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)us, R, per_real, sum_us);
  else ensemble_sum_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)us, R, per_real, sum_us);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_small_sum(smm_ctx* ctx, const double* v, int R, int per_real, double* out) {
  ensemble_sum_kernel<double><<<(per_real + 255) / 256, 256, 0, ctx->stream>>>(v, R, (size_t)per_real, out);
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
