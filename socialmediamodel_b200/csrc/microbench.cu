// microbench.cu -- measured denominators for the rooflines that MEASURED_PEAKS.json does not carry (it has the HBM copy
// rate and the bf16 GEMM rate only) and for the floor of the end-to-end ensemble call.  Each entry times a few-line
// kernel with CUDA events on the context's stream, after a warm-up launch:
//   MUFU    ex2.approx issue rate                        -> bound of the dense ABM pair loop (abm.jl:47-67: exp + sqrt per pair)
//   GATHER  random 8-byte gathers from an L2-resident table -> bound of the CSR population kernel's neighbour reads (C5)
//   LDS     shared-memory 16-byte load rate              -> bound of the FFT passes (pde.jl:49-59 evaluated by FFT)
//   D2H / WIDEN  PCIe device->host rate, host Float32 -> Float64 streaming rate: floor of smm_abm_ensemble with host buffers
#include <stdlib.h>

#include <algorithm>

#include "host_widen.h"
#include "smm_internal.h"

namespace {

// OP: 0 ex2, 1 sqrt, 2 rsqrt, 3 sqrt then ex2 (the pair loop's chain), 4 rsqrt then ex2, 5 lg2, 6 rcp, 7 sin
template <int OP>
__device__ __forceinline__ void mufu_op(float& v) {
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v));
  if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v));
  if (OP == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v));
  if (OP == 3) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v)); }
  if (OP == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v)); }
  if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v));
  if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v));
  if (OP == 7) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(v));
}
template <int OP>
__global__ void __launch_bounds__(256) mufu_kernel(int iters, float* sink) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = (OP == 0 ? -1.0f : 1.0f) * (1.0f + 0.01f * (float)(threadIdx.x + k));
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) mufu_op<OP>(v[k]);  // 8 independent chains
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += v[k];
  if (s == 123.456f) sink[0] = s;
}
template <int OP>
void launch_mufu(int grid, int iters, float* sink, cudaStream_t st) { mufu_kernel<OP><<<grid, 256, 0, st>>>(iters, sink); }
typedef void (*mufu_launcher)(int, int, float*, cudaStream_t);
const mufu_launcher kMufu[8] = {launch_mufu<0>, launch_mufu<1>, launch_mufu<2>, launch_mufu<3>, launch_mufu<4>, launch_mufu<5>, launch_mufu<6>, launch_mufu<7>};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void fill_index_kernel(int32_t* idx, size_t count, uint32_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
    idx[e] = (int32_t)(hash32((uint32_t)e * 2654435761u + 12345u) % n);
}
// thread = one "agent": deg consecutive indices (streamed, like a CSR row), deg random 8-byte gathers, one 8-byte store
template <int DEG>
__global__ void __launch_bounds__(256) gather_kernel(const float2* __restrict__ tab, const int32_t* __restrict__ idx, float2* __restrict__ out, uint32_t n) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t* row = idx + (size_t)j * DEG;
  float sx = 0, sy = 0;
#pragma unroll
  for (int e = 0; e < DEG; e++) {
    const float2 v = tab[row[e]];
    sx += v.x; sy += v.y;
  }
  out[j] = make_float2(sx, sy);
}

__global__ void __launch_bounds__(512) lds_kernel(int iters, float* sink) {
  extern __shared__ float4 sh4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh4[i] = make_float4(1.f, 2.f, 3.f, (float)i);
  __syncthreads();
  float4 a = make_float4(0, 0, 0, 0);
  int p = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {  // conflict-free: a warp reads 512 contiguous bytes
      const float4 v = sh4[(p + k * 512) & 4095];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    p = (p + 32) & 4095;
  }
  if (a.x + a.y + a.z + a.w == 123.456f) sink[0] = a.x;
}

}  // namespace

static int time_ms(smm_ctx* ctx, cudaEvent_t e0, cudaEvent_t e1, float* ms) {
  SMM_CUDA(ctx, cudaEventSynchronize(e1));
  SMM_CUDA(ctx, cudaEventElapsedTime(ms, e0, e1));
  return SMM_OK;
}

extern "C" int smm_microbench(smm_ctx* ctx, int32_t which, uint64_t size, double* result) {
  SMM_REQUIRE(ctx, ctx && result, "smm_microbench: NULL argument");
  SMM_CUDA(ctx, cudaSetDevice(ctx->device));
  *result = 0.0;
  cudaEvent_t e0, e1;
  SMM_CUDA(ctx, cudaEventCreate(&e0));
  SMM_CUDA(ctx, cudaEventCreate(&e1));
  int rc = SMM_OK;
  float ms = 0.f;
  void* sink = nullptr;
  rc = smm_scratch(ctx, S_TMP4, 256, &sink);
  if (rc == SMM_OK && which == SMM_MB_MUFU) {
    const int iters = 4096, grid = ctx->sm_count * 8;
    const int op = size < 8 ? (int)size : 0;  // size selects the operation (0: ex2, the default; see mufu_op)
    kMufu[op](grid, 64, (float*)sink, ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    kMufu[op](grid, iters, (float*)sink, ctx->stream);
    cudaEventRecord(e1, ctx->stream);
    ctx->launches += 2;
    const int per = (op == 3 || op == 4) ? 2 : 1;
    if ((rc = time_ms(ctx, e0, e1, &ms)) == SMM_OK) *result = (double)grid * 256 * 8 * iters * per / (ms * 1e-3) / 1e9;  // G MUFU/s
  } else if (rc == SMM_OK && which == SMM_MB_LDS) {
    const int iters = 2048, grid = ctx->sm_count * 2;
    cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    lds_kernel<<<grid, 512, 65536, ctx->stream>>>(16, (float*)sink);
    cudaEventRecord(e0, ctx->stream);
    lds_kernel<<<grid, 512, 65536, ctx->stream>>>(iters, (float*)sink);
    cudaEventRecord(e1, ctx->stream);
    ctx->launches += 2;
    if ((rc = time_ms(ctx, e0, e1, &ms)) == SMM_OK) *result = (double)grid * 512 * 8 * iters * 16 / (ms * 1e-3) / 1e9;  // GB/s
  } else if (rc == SMM_OK && which == SMM_MB_GATHER) {
    // size = table entries (C5: 1e7 float2 = 80 MB); 17 gathers per thread, persisting-L2 window over the table as in
    // smm_launch_pop_step
    const uint32_t n = (uint32_t)std::max<uint64_t>(size, 1024);
    constexpr int DEG = 17;
    void *tab = nullptr, *idx = nullptr, *out = nullptr;
    if ((rc = smm_scratch(ctx, S_TMP0, (size_t)n * 8, &tab)) == SMM_OK && (rc = smm_scratch(ctx, S_TMP1, (size_t)n * DEG * 4, &idx)) == SMM_OK &&
        (rc = smm_scratch(ctx, S_TMP2, (size_t)n * 8, &out)) == SMM_OK) {
      cudaMemsetAsync(tab, 0, (size_t)n * 8, ctx->stream);
      fill_index_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((int32_t*)idx, (size_t)n * DEG, n);
      cudaDeviceProp prop;
      size_t saved = 0;
      bool window = false;
      if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
          cudaDeviceGetLimit(&saved, cudaLimitPersistingL2CacheSize) == cudaSuccess) {
        const size_t want = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)n * 8);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::max(saved, want)) == cudaSuccess) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof av);
          const size_t win = std::min((size_t)n * 8, (size_t)prop.accessPolicyMaxWindowSize);
          av.accessPolicyWindow.base_ptr = tab;
          av.accessPolicyWindow.num_bytes = win;
          av.accessPolicyWindow.hitRatio = (float)(want >= win ? 1.0 : (double)want / (double)win);
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
          window = true;
        }
      }
      const unsigned grid = (n + 255) / 256;
      gather_kernel<DEG><<<grid, 256, 0, ctx->stream>>>((const float2*)tab, (const int32_t*)idx, (float2*)out, n);
      gather_kernel<DEG><<<grid, 256, 0, ctx->stream>>>((const float2*)tab, (const int32_t*)idx, (float2*)out, n);
      cudaEventRecord(e0, ctx->stream);
      for (int r = 0; r < 5; r++) gather_kernel<DEG><<<grid, 256, 0, ctx->stream>>>((const float2*)tab, (const int32_t*)idx, (float2*)out, n);
      cudaEventRecord(e1, ctx->stream);
      ctx->launches += 8;
      if ((rc = time_ms(ctx, e0, e1, &ms)) == SMM_OK) *result = (double)n * DEG * 5 / (ms * 1e-3) / 1e9;  // G gathers/s
      if (window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaStreamSynchronize(ctx->stream);
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, saved);
      }
    }
  } else if (rc == SMM_OK && (which == SMM_MB_D2H || which == SMM_MB_WIDEN)) {
    const size_t bytes = std::max<uint64_t>(size, 1 << 20), nfl = bytes / 4;
    void *dev = nullptr, *pin = nullptr;
    if ((rc = smm_scratch(ctx, S_TMP0, nfl * 4, &dev)) == SMM_OK) {
      if (cudaHostAlloc(&pin, nfl * 4, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); rc = smm_fail(ctx, SMM_ERR_NOMEM, "smm_microbench: pinned allocation failed"); }
    }
    if (rc == SMM_OK && which == SMM_MB_D2H) {
      cudaMemcpyAsync(pin, dev, nfl * 4, cudaMemcpyDeviceToHost, ctx->stream);
      cudaEventRecord(e0, ctx->stream);
      for (int r = 0; r < 3; r++) cudaMemcpyAsync(pin, dev, nfl * 4, cudaMemcpyDeviceToHost, ctx->stream);
      cudaEventRecord(e1, ctx->stream);
      if ((rc = time_ms(ctx, e0, e1, &ms)) == SMM_OK) *result = (double)nfl * 4 * 3 / (ms * 1e-3) / 1e9;  // GB/s over PCIe
    } else if (rc == SMM_OK) {
      double* dst = (double*)malloc(nfl * 8);  // pageable, like the caller's Julia array
      if (!dst) rc = smm_fail(ctx, SMM_ERR_NOMEM, "smm_microbench: host allocation failed");
      else {
        memset(pin, 0, nfl * 4);
        if (!ctx->hostpool) ctx->hostpool = smm_hostpool_create(ctx->host_threads);
        smm_hostpool_widen(ctx->hostpool, (const float*)pin, dst, nfl);  // touches the pages
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int r = 0; r < 3; r++) smm_hostpool_widen(ctx->hostpool, (const float*)pin, dst, nfl);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        *result = (double)nfl * 8 * 3 / s / 1e9;  // GB/s of Float64 written
        free(dst);
      }
    }
    if (pin) cudaFreeHost(pin);
  } else if (rc == SMM_OK) {
    rc = smm_fail(ctx, SMM_ERR_INVALID, "smm_microbench: unknown benchmark %d", which);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc == SMM_OK) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) rc = smm_fail(ctx, SMM_ERR_CUDA, "smm_microbench: %s", cudaGetErrorString(e)); }
  return rc;
}
