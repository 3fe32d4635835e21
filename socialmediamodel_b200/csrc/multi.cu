// multi.cu -- several GPUs behind ONE context and ONE call (threading row: "multi-GPU is inside one
// call: single process, ncclCommInitAll").  A Julia user calls ABMsolveensemble(200, 1024) from one task; the
// realizations (ensemble.jl:24-25) / PDE instances (ensemble.jl:69-70) are independent, so the library shards
// them over the member GPUs (one host thread per GPU, each driving an ordinary single-GPU context), every GPU
// writes its slice of us / zs / ys straight into the caller's arrays, and the only exchange is one NCCL
// all-reduce of the ensemble sums (plotting.jl:361-363 numerators + the follower numbers, plotting.jl:284-307).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 — the copy torch already loaded when there is one), so the
// library has no link-time dependency on it; without NCCL the sums are reduced on the host and the context says
// so (smm_ctx_nccl_active() == 0).
#include <dlfcn.h>
#include <nccl.h>  // types only; every entry point goes through dlsym

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "smm_internal.h"

// One persistent host thread per member GPU (a fresh std::thread per call pays the CUDA runtime's per-thread
// device binding again every time: measured 14 ms per 2-GPU call).
struct smm_worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = true, quit = false;
  void loop() {
    for (;;) {
      std::function<void()> j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_job || quit; });
        if (quit) return;
        j = std::move(job);
        has_job = false;
      }
      j();
      {
        std::lock_guard<std::mutex> lk(mu);
        done = true;
      }
      cv.notify_all();
    }
  }
  void start() { th = std::thread([this] { loop(); }); }
  void submit(std::function<void()> j) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = std::move(j);
      has_job = true;
      done = false;
    }
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
    }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

struct smm_multi {
  int n = 0;
  std::vector<smm_worker*> workers;  // workers[g] drives member g, g >= 1 (member 0 runs on the calling thread)
  std::vector<int> devices;
  std::vector<smm_ctx*> member;
  // NCCL (single process, one communicator per member GPU)
  void* dl = nullptr;
  std::vector<ncclComm_t> comms;
  bool nccl = false;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  int64_t allreduces = 0;
};

static void multi_free(smm_multi* m) {
  if (!m) return;
  for (smm_worker* w : m->workers) if (w) { w->stop(); delete w; }
  if (m->nccl)
    for (ncclComm_t c : m->comms) if (c) m->CommDestroy(c);
  for (smm_ctx* c : m->member) smm_ctx_destroy(c);
  if (m->dl) dlclose(m->dl);
  delete m;
}

void smm_multi_destroy(smm_multi* m) { multi_free(m); }

static bool multi_init_nccl(smm_multi* m) {
  const char* off = getenv("SMM_MULTI_NCCL");
  if (off && off[0] == '0') return false;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    m->dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (m->dl) break;
  }
  if (!m->dl) return false;
#define SYM(field, name) *(void**)(&m->field) = dlsym(m->dl, name)
  SYM(CommInitAll, "ncclCommInitAll"); SYM(CommDestroy, "ncclCommDestroy"); SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  if (!m->CommInitAll || !m->CommDestroy || !m->AllReduce || !m->GroupStart || !m->GroupEnd) return false;
  m->comms.assign(m->n, nullptr);
  if (m->CommInitAll(m->comms.data(), m->n, m->devices.data()) != ncclSuccess) {
    m->comms.clear();
    cudaGetLastError();
    return false;
  }
  return true;
}

extern "C" int smm_ctx_create_multi(int ngpus, const int* device_ids, smm_ctx** out) {
  if (!out) return smm_fail(nullptr, SMM_ERR_INVALID, "smm_ctx_create_multi: out is NULL");
  *out = nullptr;
  const int have = smm_device_count();
  if (have == 0) return smm_fail(nullptr, SMM_ERR_CUDA, "smm_ctx_create_multi: no CUDA device; libsmm_b200 has no CPU fallback");
  if (ngpus <= 0) ngpus = have;
  if (ngpus > have && !device_ids) return smm_fail(nullptr, SMM_ERR_INVALID, "smm_ctx_create_multi: %d GPUs requested, %d visible", ngpus, have);
  smm_multi* m = new smm_multi();
  m->n = ngpus;
  for (int g = 0; g < ngpus; g++) {
    const int d = device_ids ? device_ids[g] : g;
    for (int h = 0; h < g; h++)
      if (m->devices[h] == d) { multi_free(m); return smm_fail(nullptr, SMM_ERR_INVALID, "smm_ctx_create_multi: device %d listed twice", d); }
    m->devices.push_back(d);
  }
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 4;
  for (int g = 0; g < ngpus; g++) {
    smm_ctx* c = nullptr;
    int rc = smm_ctx_create(m->devices[g], &c);
    if (rc != SMM_OK) { multi_free(m); return rc; }
    c->host_threads = std::max(1, std::min(16, hw / ngpus));  // host widening threads of this member (host_widen.cpp)
    c->shard_mode = true;
    m->member.push_back(c);
  }
  m->workers.assign(ngpus, nullptr);
  for (int g = 1; g < ngpus; g++) { m->workers[g] = new smm_worker(); m->workers[g]->start(); }
  m->nccl = ngpus > 1 && multi_init_nccl(m);
  smm_ctx* lead = nullptr;
  int rc = smm_ctx_create(m->devices[0], &lead);
  if (rc != SMM_OK) { multi_free(m); return rc; }
  lead->multi = m;
  *out = lead;
  return SMM_OK;
}

extern "C" int smm_ctx_gpu_count(const smm_ctx* ctx) { return ctx ? (ctx->multi ? ctx->multi->n : 1) : 0; }
extern "C" int smm_ctx_nccl_active(const smm_ctx* ctx) { return (ctx && ctx->multi && ctx->multi->nccl) ? 1 : 0; }

int64_t smm_multi_launches(const smm_multi* m) {
  int64_t s = 0;
  for (const smm_ctx* c : m->member) s += c->launches;
  return s;
}
double smm_multi_kernel_ms(const smm_multi* m, int which) {  // slowest member: the call waits for it
  double v = -1.0;
  for (const smm_ctx* c : m->member) v = std::max(v, c->kernel_ms[which]);
  return v;
}

static void shard(int64_t N, int g, int n, int64_t* lo, int64_t* hi) {  // contiguous blocks, sizes differ by <= 1 (dist.shard_range)
  const int64_t base = N / n, rem = N % n;
  *lo = g * base + std::min<int64_t>(g, rem);
  *hi = *lo + base + (g < rem ? 1 : 0);
}

// Sum the member GPUs' device buffers `bufs[g]` (count elements each; members with no work hold zeros) into every
// member's buffer: one grouped ncclAllReduce over NVLink, or nothing when NCCL is unavailable (the caller then
// adds the per-member downloads on the host).
static int multi_allreduce(smm_ctx* lead, std::vector<void*>& bufs, size_t count, ncclDataType_t type) {
  smm_multi* m = lead->multi;
  if (!m->nccl) return SMM_OK;
  ncclResult_t r = m->GroupStart();
  for (int g = 0; g < m->n && r == ncclSuccess; g++)
    r = m->AllReduce(bufs[g], bufs[g], count, type, ncclSum, m->comms[g], m->member[g]->stream);
  ncclResult_t r2 = m->GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess)
    return smm_fail(lead, SMM_ERR_CUDA, "ncclAllReduce of the ensemble sums failed: %s", m->GetErrorString ? m->GetErrorString(r) : "?");
  m->allreduces++;
  return SMM_OK;
}

int smm_multi_abm_ensemble(smm_ctx* lead, const smm_params* q, int32_t precision, int32_t NT, int32_t R, int64_t first,
                           const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0, const double* media0,
                           const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                           const int32_t* savesteps, double kde_sigma, const smm_grid* grid, void* us, bool us_f32, double* zs,
                           double* ys, double* sum_us, double* sum_zs, double* sum_ys, int64_t* numbers) {
  smm_multi* m = lead->multi;
  SMM_REQUIRE(lead, q && noise && grid, "smm_abm_ensemble (multi-GPU): NULL argument");
  SMM_REQUIRE(lead, R >= 1 && NT >= 0 && nsave >= 0, "smm_abm_ensemble (multi-GPU): need R >= 1, NT >= 0, nsave >= 0");
  const size_t n = q->n, L = q->L, M = q->M;
  const size_t per_real = (size_t)grid->Nx * grid->Ny * 2 * L * nsave;
  const size_t us_elem = us_f32 ? 4 : 8;
  const int G = m->n;
  std::vector<int> rcs(G, SMM_OK);
  std::vector<int64_t> lo(G), hi(G);
  for (int g = 0; g < G; g++) shard(R, g, G, &lo[g], &hi[g]);
  const bool want_sums = (sum_us || sum_zs || sum_ys || numbers);
  // per-member host landing buffers for the sums when NCCL is unavailable
  std::vector<std::vector<double>> h_us(G), h_zs(G), h_ys(G);
  std::vector<std::vector<int64_t>> h_num(G);
  const bool host_reduce = want_sums && !m->nccl;
  auto work = [&](int g) {
    smm_ctx* c = m->member[g];
    const int64_t r0 = lo[g], rc = hi[g] - lo[g];
    if (rc <= 0) return;
    smm_noise ns = *noise;
    if (ns.mode == SMM_NOISE_INJECTED) {
      if (ns.xi) ns.xi += (size_t)r0 * NT * 2 * n;
      if (ns.u1) ns.u1 += (size_t)r0 * NT * n;
      if (ns.u2) ns.u2 += (size_t)r0 * NT * n;
      if (ns.xi_inf) ns.xi_inf += (size_t)r0 * NT * L * 2;
      if (ns.xi_med) ns.xi_med += (size_t)r0 * NT * M * 2;
    }
    double *su = nullptr, *sz = nullptr, *sy = nullptr;
    int64_t* nu = nullptr;
    if (host_reduce) {
      if (sum_us) { h_us[g].assign(per_real, 0.0); su = h_us[g].data(); }
      if (sum_zs) { h_zs[g].assign((size_t)nsave * M * 2, 0.0); sz = h_zs[g].data(); }
      if (sum_ys) { h_ys[g].assign((size_t)nsave * L * 2, 0.0); sy = h_ys[g].data(); }
      if (numbers) { h_num[g].assign((size_t)(NT + 1) * L * M, 0); nu = h_num[g].data(); }
    } else {  // NCCL: the sums stay in the member's device scratch (shard_mode skips their download)
      su = sum_us; sz = sum_zs; sy = sum_ys; nu = numbers;
    }
    c->shard_mode = !host_reduce;
    rcs[g] = smm_abm_ensemble_shard(
        c, q, precision, NT, (int32_t)rc, first + r0, x0 ? x0 + (size_t)r0 * 2 * n : nullptr, state ? state + (size_t)r0 * n : nullptr,
        follow0 ? follow0 + (size_t)r0 * n : nullptr, inf0 ? inf0 + (size_t)r0 * 2 * L : nullptr,
        media0 ? media0 + (size_t)r0 * 2 * M : nullptr, rowptr, colidx, &ns, nsave, savesteps, kde_sigma, grid,
        us ? (char*)us + (size_t)r0 * per_real * us_elem : nullptr, us_f32, zs ? zs + (size_t)r0 * nsave * M * 2 : nullptr,
        ys ? ys + (size_t)r0 * nsave * L * 2 : nullptr, su, sz, sy, nu);
  };
  for (int g = 1; g < G; g++) m->workers[g]->submit([&work, g] { work(g); });
  work(0);
  for (int g = 1; g < G; g++) m->workers[g]->wait();
  for (int g = 0; g < G; g++)
    if (rcs[g] != SMM_OK) return smm_fail(lead, rcs[g], "GPU %d: %s", m->devices[g], m->member[g]->err.c_str());
  lead->launches = smm_multi_launches(m);
  for (int k = 0; k < SMM_K_NCLASS; k++) lead->kernel_ms[k] = smm_multi_kernel_ms(m, k);
  if (!want_sums) return SMM_OK;
  if (host_reduce) {
    auto add = [&](double* dst, std::vector<std::vector<double>>& src, size_t cnt) {
      if (!dst) return;
      std::fill(dst, dst + cnt, 0.0);
      for (int g = 0; g < G; g++)
        if (!src[g].empty()) for (size_t e = 0; e < cnt; e++) dst[e] += src[g][e];
    };
    add(sum_us, h_us, per_real); add(sum_zs, h_zs, (size_t)nsave * M * 2); add(sum_ys, h_ys, (size_t)nsave * L * 2);
    if (numbers) {
      const size_t cnt = (size_t)(NT + 1) * L * M;
      std::fill(numbers, numbers + cnt, (int64_t)0);
      for (int g = 0; g < G; g++)
        if (!h_num[g].empty()) for (size_t e = 0; e < cnt; e++) numbers[e] += h_num[g][e];
    }
    return SMM_OK;
  }
  // ---- NCCL: one all-reduce per sum buffer, in place in the members' scratch; the result is read from GPU 0
  struct Red { int slot; size_t count; ncclDataType_t type; void* host; size_t bytes; };
  std::vector<Red> reds;
  if (sum_us && nsave) reds.push_back({S_SUMUS, per_real, ncclDouble, sum_us, per_real * 8});
  if (sum_zs && nsave) reds.push_back({S_SUMZS, (size_t)nsave * M * 2, ncclDouble, sum_zs, (size_t)nsave * M * 2 * 8});
  if (sum_ys && nsave) reds.push_back({S_SUMYS, (size_t)nsave * L * 2, ncclDouble, sum_ys, (size_t)nsave * L * 2 * 8});
  if (numbers) reds.push_back({S_NUMBERS, (size_t)(NT + 1) * L * M, ncclInt64, numbers, (size_t)(NT + 1) * L * M * 8});
  for (const Red& r : reds) {
    std::vector<void*> bufs(G);
    for (int g = 0; g < G; g++) {
      smm_ctx* c = m->member[g];
      cudaSetDevice(c->device);
      void* p = nullptr;
      SMM_TRY(smm_scratch(c, r.slot, r.bytes, &p));
      if (hi[g] - lo[g] <= 0) cudaMemsetAsync(p, 0, r.bytes, c->stream);  // more GPUs than members: contribute zeros
      bufs[g] = p;
    }
    SMM_TRY(multi_allreduce(lead, bufs, r.count, r.type));
    smm_ctx* c0 = m->member[0];
    cudaSetDevice(c0->device);
    SMM_CUDA(lead, cudaMemcpyAsync(r.host, bufs[0], r.bytes, cudaMemcpyDeviceToHost, c0->stream));
  }
  for (int g = 0; g < G; g++) {
    cudaSetDevice(m->member[g]->device);
    SMM_CUDA(lead, cudaStreamSynchronize(m->member[g]->stream));
  }
  cudaSetDevice(lead->device);
  return SMM_OK;
}

// PDEsolveensemble / parameter sweep (ensemble.jl:58-85): instances sharded the same way.  The per-save-point sums are
// formed on the host by smm_pde_solve (the states are downloaded anyway), so the members' sums are added on the host.
int smm_multi_pde_solve(smm_ctx* lead, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L, int32_t precision,
                        const double* controlled_inf, const double* controlled_med, const double* u0, const double* z0,
                        const double* y0, double dt, int32_t nsteps, int32_t scheme, int32_t nsave, const int32_t* savesteps,
                        double* us, double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys) {
  smm_multi* m = lead->multi;
  SMM_REQUIRE(lead, grid && q && u0 && z0 && y0 && B >= 1, "smm_pde_solve (multi-GPU): bad argument");
  const int G = m->n;
  const size_t per = (size_t)grid->Nx * grid->Ny * 2 * L;
  std::vector<int> rcs(G, SMM_OK);
  std::vector<int64_t> lo(G), hi(G);
  for (int g = 0; g < G; g++) shard(B, g, G, &lo[g], &hi[g]);
  std::vector<std::vector<double>> h_us(G), h_zs(G), h_ys(G);
  auto work = [&](int g) {
    const int64_t b0 = lo[g], bc = hi[g] - lo[g];
    if (bc <= 0) return;
    if (sum_us) h_us[g].assign(per * nsave, 0.0);
    if (sum_zs) h_zs[g].assign((size_t)4 * nsave, 0.0);
    if (sum_ys) h_ys[g].assign((size_t)2 * L * nsave, 0.0);
    rcs[g] = smm_pde_solve(m->member[g], grid, q + b0, (int32_t)bc, L, precision, controlled_inf ? controlled_inf + (size_t)b0 * L : nullptr,
                           controlled_med ? controlled_med + (size_t)b0 * 2 : nullptr, u0 + (size_t)b0 * per, z0 + (size_t)b0 * 4,
                           y0 + (size_t)b0 * 2 * L, dt, nsteps, scheme, nsave, savesteps, us ? us + (size_t)b0 * nsave * per : nullptr,
                           zs ? zs + (size_t)b0 * nsave * 4 : nullptr, ys ? ys + (size_t)b0 * nsave * 2 * L : nullptr,
                           sum_us ? h_us[g].data() : nullptr, sum_zs ? h_zs[g].data() : nullptr, sum_ys ? h_ys[g].data() : nullptr);
  };
  for (int g = 1; g < G; g++) m->workers[g]->submit([&work, g] { work(g); });
  work(0);
  for (int g = 1; g < G; g++) m->workers[g]->wait();
  for (int g = 0; g < G; g++)
    if (rcs[g] != SMM_OK) return smm_fail(lead, rcs[g], "GPU %d: %s", m->devices[g], m->member[g]->err.c_str());
  lead->launches = smm_multi_launches(m);
  for (int k = 0; k < SMM_K_NCLASS; k++) lead->kernel_ms[k] = smm_multi_kernel_ms(m, k);
  auto add = [&](double* dst, std::vector<std::vector<double>>& src, size_t cnt) {
    if (!dst) return;
    std::fill(dst, dst + cnt, 0.0);
    for (int g = 0; g < G; g++)
      if (!src[g].empty()) for (size_t e = 0; e < cnt; e++) dst[e] += src[g][e];
  };
  add(sum_us, h_us, per * nsave); add(sum_zs, h_zs, (size_t)4 * nsave); add(sum_ys, h_ys, (size_t)2 * L * nsave);
  cudaSetDevice(lead->device);
  return SMM_OK;
}
