// mio_mix.cu -- do MUFU, SHFL and LDS instructions of one SM sub-partition overlap, or do they share one dispatch path?
// Per iteration and warp: NM independent MUFU.EX2, NS independent SHFL.IDX, NL independent LDS.128 (conflict-free).
// Prints cycles per iteration and SM sub-partition (8 warps each) -- compare (4,0,0), (0,2,0), (0,0,2) with (4,2,2).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mio_mix mio_mix.cu && ./mio_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NM, int NS, int NL, int NF>
__global__ void __launch_bounds__(256) mix_kernel(int iters, float* sink) {
  __shared__ float4 tab[512];
  for (int i = threadIdx.x; i < 512; i += 256) tab[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float m[NM > 0 ? NM : 1], s[NS > 0 ? NS : 1], f[NF > 0 ? NF : 1];
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < NM; k++) m[k] = -1.0f - 0.01f * (threadIdx.x + k);
#pragma unroll
  for (int k = 0; k < NS; k++) s[k] = threadIdx.x + k;
#pragma unroll
  for (int k = 0; k < NF; k++) f[k] = 0.5f + k;
  const float4* p = tab + lane;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < (NM > NS ? (NM > NL ? NM : NL) : (NS > NL ? NS : NL)); k++) {
      if (k < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k]));
      if (k < NS) s[k] = __shfl_sync(0xffffffffu, s[k], lane + 1 + k);
      if (k < NL) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(__cvta_generic_to_shared(p + 32 * ((k + i) & 7))));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < NF; k++) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[k]));
  }
  float r = acc.x + acc.y + acc.z + acc.w;
#pragma unroll
  for (int k = 0; k < NM; k++) r += m[k];
#pragma unroll
  for (int k = 0; k < NS; k++) r += s[k];
#pragma unroll
  for (int k = 0; k < NF; k++) r += f[k];
  if (r == 123.456f) sink[0] = r;
}

template <int NM, int NS, int NL, int NF>
void run(float* sink, int sms, double ghz) {
  const int iters = 20000, grid = sms * 4;  // 4 CTAs x 8 warps = 32 warps per SM = 8 per sub-partition
  mix_kernel<NM, NS, NL, NF><<<grid, 256>>>(100, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  mix_kernel<NM, NS, NL, NF><<<grid, 256>>>(iters, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double cyc = ms * 1e-3 * ghz * 1e9 / iters / 8.0;  // cycles per warp-iteration on one sub-partition
  printf("MUFU %d SHFL %d LDS.128 %d FMA %d : %7.2f cycles per warp-iteration and sub-partition\n", NM, NS, NL, NF, cyc);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (nominal; cycles assume this clock)\n", pr.name, pr.multiProcessorCount, ghz);
  float* sink;
  cudaMalloc(&sink, 256);
  const int sms = pr.multiProcessorCount;
  run<4, 0, 0, 0>(sink, sms, ghz);
  run<0, 2, 0, 0>(sink, sms, ghz);
  run<0, 0, 2, 0>(sink, sms, ghz);
  run<4, 2, 0, 0>(sink, sms, ghz);
  run<4, 0, 2, 0>(sink, sms, ghz);
  run<4, 2, 2, 0>(sink, sms, ghz);
  run<0, 2, 2, 0>(sink, sms, ghz);
  run<8, 6, 1, 0>(sink, sms, ghz);
  run<0, 8, 0, 0>(sink, sms, ghz);
  run<0, 0, 8, 0>(sink, sms, ghz);
  run<4, 2, 2, 11>(sink, sms, ghz);
  run<4, 0, 0, 11>(sink, sms, ghz);
  run<0, 0, 0, 16>(sink, sms, ghz);
  cudaFree(sink);
  return 0;
}
