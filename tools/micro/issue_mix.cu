// issue_mix.cu -- does a packed FP32x2 instruction (FFMA2) take one issue slot or two?  Per iteration and warp:
// NP packed FMAs (fma.rn.f32x2), NS scalar FMAs, NI integer adds (ALU pipe), NM MUFU.  8 warps per sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NP, int NS, int NI, int NM>
__global__ void __launch_bounds__(256) mix_kernel(int iters, float* sink) {
  unsigned long long p[NP > 0 ? NP : 1];
  float s[NS > 0 ? NS : 1], m[NM > 0 ? NM : 1];
  unsigned q[NI > 0 ? NI : 1];
#pragma unroll
  for (int k = 0; k < NP; k++) p[k] = 0x3f0000003f000000ull + threadIdx.x + k;
#pragma unroll
  for (int k = 0; k < NS; k++) s[k] = 0.5f + k + threadIdx.x;
#pragma unroll
  for (int k = 0; k < NI; k++) q[k] = threadIdx.x + k;
#pragma unroll
  for (int k = 0; k < NM; k++) m[k] = -1.0f - 0.01f * (threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k < NP) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[k]));
      if (k < NS) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(s[k]));
      if (k < NI) asm volatile("add.u32 %0, %0, %1;" : "+r"(q[k]) : "r"(i));
      if (k < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k]));
    }
  }
  float r = 0;
#pragma unroll
  for (int k = 0; k < NP; k++) r += (float)(p[k] & 0xffff);
#pragma unroll
  for (int k = 0; k < NS; k++) r += s[k];
#pragma unroll
  for (int k = 0; k < NI; k++) r += (float)q[k];
#pragma unroll
  for (int k = 0; k < NM; k++) r += m[k];
  if (r == 123.456f) sink[0] = r;
}

template <int NP, int NS, int NI, int NM>
void run(float* sink, int sms, double ghz) {
  const int iters = 20000, grid = sms * 4;
  mix_kernel<NP, NS, NI, NM><<<grid, 256>>>(100, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  mix_kernel<NP, NS, NI, NM><<<grid, 256>>>(iters, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("FFMA2 %2d FFMA %2d IADD %2d MUFU %2d : %7.2f cycles per warp-iteration and sub-partition\n", NP, NS, NI, NM,
         ms * 1e-3 * ghz * 1e9 / iters / 8.0);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float* sink;
  cudaMalloc(&sink, 256);
  const int sms = pr.multiProcessorCount;
  run<16, 0, 0, 0>(sink, sms, ghz);
  run<0, 16, 0, 0>(sink, sms, ghz);
  run<0, 0, 16, 0>(sink, sms, ghz);
  run<16, 0, 16, 0>(sink, sms, ghz);
  run<8, 0, 16, 0>(sink, sms, ghz);
  run<0, 16, 16, 0>(sink, sms, ghz);
  run<8, 8, 0, 0>(sink, sms, ghz);
  run<0, 0, 16, 4>(sink, sms, ghz);
  run<0, 16, 0, 4>(sink, sms, ghz);
  run<10, 0, 4, 4>(sink, sms, ghz);
  run<10, 0, 8, 4>(sink, sms, ghz);
  cudaFree(sink);
  return 0;
}
