// host_widen.cpp — host side of the ensemble output pipeline.
//
// smm_abm_ensemble returns the per-realization densities `us` as Float64 in the caller's (Julia) array,
// `ensemble.jl:20,38` — 2 GB for 1 024 realizations at 89x89.  In FP32 mode the device computes them in
// FP32; widening them on the GPU doubles the PCIe traffic, which is what bounds the end-to-end call.
// Instead the FP32 chunks cross PCIe into a pinned ring and a small pool of host threads widens them
// (exactly: every float is representable as a double) into the destination with streaming stores,
// overlapped with the DMA of the next chunk and the simulation of the one after.  The destination can
// be ordinary pageable memory.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>
#include <immintrin.h>

#include "host_widen.h"

namespace {

void widen_scalar(const float* src, double* dst, size_t count) {
  for (size_t i = 0; i < count; i++) dst[i] = (double)src[i];
}

__attribute__((target("avx2"))) void widen_avx2(const float* src, double* dst, size_t count) {
  size_t i = 0;
  while (i < count && ((uintptr_t)(dst + i) & 31)) { dst[i] = (double)src[i]; i++; }
  for (; i + 8 <= count; i += 8) {
    const __m128 lo = _mm_loadu_ps(src + i), hi = _mm_loadu_ps(src + i + 4);
    _mm256_stream_pd(dst + i, _mm256_cvtps_pd(lo));
    _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(hi));
  }
  for (; i < count; i++) dst[i] = (double)src[i];
  _mm_sfence();
}

// one 64-byte streaming store per destination cache line (AVX2 needs two 32-byte halves, which the write-combining buffers
// have to merge)
__attribute__((target("avx512f"))) void widen_avx512(const float* src, double* dst, size_t count) {
  size_t i = 0;
  while (i < count && ((uintptr_t)(dst + i) & 63)) { dst[i] = (double)src[i]; i++; }
  for (; i + 16 <= count; i += 16) {
    const __m256 lo = _mm256_loadu_ps(src + i), hi = _mm256_loadu_ps(src + i + 8);
    _mm512_stream_pd(dst + i, _mm512_cvtps_pd(lo));
    _mm512_stream_pd(dst + i + 8, _mm512_cvtps_pd(hi));
  }
  for (; i < count; i++) dst[i] = (double)src[i];
  _mm_sfence();
}

}  // namespace

struct smm_hostpool {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  uint64_t generation = 0;
  int remaining = 0;
  bool quit = false;
  bool avx2 = false, avx512 = false;
  const float* src = nullptr;
  double* dst = nullptr;
  size_t count = 0;
  int parts = 1;

  void run_part(int part) const {
    // slices are multiples of 16 elements so that every slice but the first starts 32-byte aligned
    // relative to the previous one
    const size_t per = ((count + parts - 1) / parts + 15) & ~(size_t)15;
    const size_t b = std::min(count, per * (size_t)part), e = std::min(count, b + per);
    if (e <= b) return;
    if (avx512) widen_avx512(src + b, dst + b, e - b);
    else if (avx2) widen_avx2(src + b, dst + b, e - b);
    else widen_scalar(src + b, dst + b, e - b);
  }

  void worker(int part) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_start.wait(lk, [&] { return quit || generation != seen; });
        if (quit) return;
        seen = generation;
      }
      run_part(part);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--remaining == 0) cv_done.notify_one();
      }
    }
  }
};

smm_hostpool* smm_hostpool_create(int nthreads) {
  if (nthreads <= 0) {
    // default: the cores this process may reasonably use — all of them, divided by the number of
    // ranks torchrun started on this node, at most 16
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 4;
    int ranks = 1;
    if (const char* s = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(s));
    nthreads = std::max(1, std::min(16, hw / ranks));
    if (const char* s = getenv("SMM_HOST_THREADS")) nthreads = std::max(1, atoi(s));
  }
  smm_hostpool* p = new smm_hostpool;
  p->parts = nthreads;
  p->avx2 = __builtin_cpu_supports("avx2");
  {  // SMM_HOST_WIDEN_ISA=avx2 keeps the 32-byte stores (A/B timing)
    const char* isa = getenv("SMM_HOST_WIDEN_ISA");
    p->avx512 = __builtin_cpu_supports("avx512f") && !(isa && isa[0] == 'a' && isa[3] == '2');
  }
  for (int t = 1; t < nthreads; t++) p->workers.emplace_back([p, t] { p->worker(t); });
  return p;
}

void smm_hostpool_destroy(smm_hostpool* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->quit = true;
  }
  p->cv_start.notify_all();
  for (auto& t : p->workers) t.join();
  delete p;
}

int smm_hostpool_threads(const smm_hostpool* p) { return p ? p->parts : 0; }

void smm_hostpool_widen(smm_hostpool* p, const float* src, double* dst, size_t count) {
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->src = src; p->dst = dst; p->count = count;
    p->remaining = (int)p->workers.size();
    p->generation++;
  }
  p->cv_start.notify_all();
  p->run_part(0);  // the calling thread takes the first slice
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv_done.wait(lk, [&] { return p->remaining == 0; });
}
