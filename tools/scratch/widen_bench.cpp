// micro-benchmark of the host widening pool (tools/scratch; not part of the library)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <initializer_list>
#include "../../socialmediamodel_b200/csrc/host_widen.h"
int main() {
  const size_t count = (size_t)260 << 20;  // 1.04 GB of floats
  float* src = (float*)aligned_alloc(64, count * 4);
  double* dst = (double*)aligned_alloc(64, count * 8);
  for (size_t i = 0; i < count; i++) src[i] = (float)(i & 1023) * 0.5f;
  memset(dst, 0, count * 8);
  printf("hardware_concurrency=%u\n", std::thread::hardware_concurrency());
  for (int T : {1, 4, 8, 16, 32}) {
    smm_hostpool* p = smm_hostpool_create(T);
    double best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      auto t0 = std::chrono::steady_clock::now();
      smm_hostpool_widen(p, src, dst, count);
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    bool ok = dst[12345] == (double)src[12345] && dst[count - 1] == (double)src[count - 1];
    printf("T=%d  %.2f ms  out %.1f GB/s  ok=%d\n", T, best, count * 8 / best * 1e-6, (int)ok);
    smm_hostpool_destroy(p);
  }
  return 0;
}
