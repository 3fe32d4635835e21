// abm_dense.cu -- K1: the ABM time-stepper for ensembles of small all-to-all populations
// (configs C1/C2: n = 250, L = 4, M = 2).  One CTA owns one realization for the whole run:
// positions, follower ids and media states live in shared memory, the NT-step loop runs
// inside the kernel, and only the requested snapshots / histories go to HBM.
//
// Replaces the loop body abm.jl:166-213 (attraction :47-67, influence :69-88, Euler-Maruyama
// :173, influencer/media update :175-198, reflection :139-150, changeinfluencer :90-137).
// Bound: MUFU (2 per ordered pair: sqrt.approx + ex2.approx), not HBM -- see DESIGN.md.
#include <algorithm>

#include "abm_device.cuh"
#include "smm_internal.h"

namespace {

constexpr int kThreads = 128;  // CTA width of the full-GPU configuration (and of the init kernel)

// 2^(-sqrt(r2)) for coordinates pre-scaled by log2(e): exp(-|x_i - x_j|) with two MUFU ops and nothing else
__device__ __forceinline__ float w_of_scaled_r2(float r2) {
  float r, w;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(-r));
  return w;
}
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// SYM: every unordered agent pair is evaluated once (w_ij = w_ji, abm.jl:59) — see phase P3.
constexpr int kSymMaxGroups = 16;  // n <= 512: the per-offset partial sums must fit shared memory
constexpr float kPadPos = 1e15f;   // padding agents of the last 32-group: exp(-|x - pad|) = 0 exactly

// ---- KDE of a snapshot inside the step kernel (same arithmetic, order and results as kde_kernel in kde.cu) ----------
constexpr int kKdeSpan = 96;   // grid columns / rows of one pass (lane owns columns l, l + 32, l + 64; a warp 12 rows)
constexpr int kKdeChunk = 16;  // agents staged per chunk: 2 x 16 x 96 floats = 12 KB in the partial-sum area
constexpr int kKdeRows = 12;
__device__ __forceinline__ float kde_gauss1d(float d, float inv2s2) { return __expf(-d * d * inv2s2); }

// ---- symmetric pair tiles ---------------------------------------------------------------------------------------
// Table of a 32-group: T[i] = log2(e) * (x_a, x_b, y_a, y_b) of the agents a = i & 31, b = (i + 16) & 31, i < 64
// (every entry twice, so that the rotated index lane +- t needs no wrap-around arithmetic).
// Measured and rejected (round 2): a table in which lanes l and l + 16 read the same entry halves the shared-memory
// wavefronts of the loads (the data pipe is the busiest unit, 75 %), but routing the weights then costs four selects
// and a mask per step: 21 -> 26 instructions per step, kernel 5.90 -> 6.40 ms.  The loop is bound by issue slots and
// per-warp latency (0.6 instructions per cycle and scheduler with 5.7 resident warps), so what pays is instructions:
// the column-side fix-up below (21 -> 19 per step) and the symmetric diagonal tile: 5.90 -> 5.60 ms.
constexpr int kTab = 64;   // table entries per 32-group
constexpr int kSchedMax = 32;  // wide CTAs: bytes of a warp's tile list (count + up to 31 tiles; 136 tiles over >= 8 warps)
struct Sums2 { float2 ws, ax, ay; };
__device__ __forceinline__ void zero(Sums2& s) { s.ws = s.ax = s.ay = make_float2(0.f, 0.f); }
// (sum of the two halves; forces scaled back by ln 2) -- explicit _rn operations: no contraction into an FMA
__device__ __forceinline__ float fold_w(const Sums2& s) { return __fadd_rn(s.ws.x, s.ws.y); }
__device__ __forceinline__ float fold_x(const Sums2& s) { return __fmul_rn(__fadd_rn(s.ax.x, s.ax.y), kLn2); }
__device__ __forceinline__ float fold_y(const Sums2& s) { return __fmul_rn(__fadd_rn(s.ay.x, s.ay.y), kLn2); }

// ordered step: row agent against the two agents of entry p (no exchange)
__device__ __forceinline__ void ordered_step(const float4 p, const float2 nrx, const float2 nry, Sums2& R) {
  const float2 dx = add2(make_float2(p.x, p.y), nrx), dy = add2(make_float2(p.z, p.w), nry);
  const float2 r2 = fma2(dx, dx, mul2(dy, dy));
  const float2 w = make_float2(w_of_scaled_r2(r2.x), w_of_scaled_r2(r2.y));
  R.ws = add2(R.ws, w); R.ax = fma2(w, dx, R.ax); R.ay = fma2(w, dy, R.ay);
}
// symmetric steps T0..T1 of the tile (row group, column group): at step t row agent l (= lane) meets the column agents
// (l + t) & 31 and (l + t + 16) & 31; the weights go to the row sums R and, through one shuffle each, to the lanes that
// own those column agents, which pair them with the rows (l - t) & 31 and (l - t - 16) & 31 -- stored next to each other
// at index l - t + 32.  The column side accumulates sum_r w x_r and subtracts x_c sum_r w once at the end (ncx = -x_c):
// two instructions fewer per step; the cancellation costs < 1e-6 of the force (|x| <= 3.2 in these units), i.e.
// < 1e-8 of a position per step.
template <int T0, int T1>
__device__ __forceinline__ void sym_steps(const float4* rowp, const float4* colp, const int lane, const float2 nrx, const float2 nry,
                                          const float2 ncx, const float2 ncy, Sums2& R, Sums2& C) {
#pragma unroll
  for (int t = T0; t <= T1; t++) {
    const float4 p = colp[t];
    const float2 dx = add2(make_float2(p.x, p.y), nrx), dy = add2(make_float2(p.z, p.w), nry);
    const float2 r2 = fma2(dx, dx, mul2(dy, dy));
    const float2 w = make_float2(w_of_scaled_r2(r2.x), w_of_scaled_r2(r2.y));
    R.ws = add2(R.ws, w); R.ax = fma2(w, dx, R.ax); R.ay = fma2(w, dy, R.ay);
    float2 wv;
    wv.x = __shfl_sync(0xffffffffu, w.x, lane - t);
    wv.y = __shfl_sync(0xffffffffu, w.y, lane - t - 16);
    const float4 q = rowp[32 - t];
    C.ws = add2(C.ws, wv); C.ax = fma2(wv, make_float2(q.x, q.y), C.ax); C.ay = fma2(wv, make_float2(q.z, q.w), C.ay);
  }
  C.ax = fma2(C.ws, ncx, C.ax); C.ay = fma2(C.ws, ncy, C.ay);
}
// diagonal tile: rotation steps 1..7 symmetric (offsets 1..7 and 17..23 = -15..-9), steps 0 and 8 ordered (step 0 holds
// the agent itself, abm.jl:42, and its opposite; the pairs at offsets 8 and 24 are met from both sides): 9 steps, not 16
__device__ __forceinline__ void diag_tile(const float4* rowp, const int lane, const float2 nrx, const float2 nry, Sums2& R) {
  Sums2 C;
  zero(R); zero(C);
  ordered_step(rowp[0], nrx, nry, R);
  sym_steps<1, 7>(rowp, rowp, lane, nrx, nry, nrx, nry, R, C);
  ordered_step(rowp[8], nrx, nry, R);
  R.ws = add2(R.ws, C.ws); R.ax = add2(R.ax, C.ax); R.ay = add2(R.ay, C.ay);
}

// THREADS = CTA width: 128 when the ensemble alone fills the GPU (1 024 realizations: 7 CTAs per SM, all resident);
// 256 / 512 for small ensembles (configs[1] at 8 GPUs = 128 realizations per GPU; C1 = one realization), where one
// realization is spread over 8 / 16 warps instead of leaving most of the SMs' issue slots idle.
template <typename real, bool SYM, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? (SYM ? 7 : 8) : (THREADS == 256 ? 3 : 1)) abm_dense_kernel(AbmDenseArgs A) {
  using real2 = typename vec2<real>::type;
  constexpr int kThreads = THREADS;  // shadows the namespace constant inside this kernel
  const AbmConst& K = A.k;
  const int n = K.n, L = K.L, M = K.M, NT = A.NT;
  // Step chunks (A.chunk_steps > 0): CTA (chunk, r) = blockIdx.x = chunk * R + r runs steps [k0, k1) of realization r and
  // hands the state to CTA (chunk + 1, r) through global memory.  Why: the warp schedulers favour the CTAs that were
  // launched first (measured: of the 7 CTAs of an SM the first finishes 200 steps after 4.1 ms, the last after 5.9 ms), so a
  // one-CTA-per-realization grid ends in a long tail at falling occupancy.  With chunks, the slot of a CTA that finishes
  // early is refilled at once by the next chunk of an early realization; the SMs stay full until the last chunk.
  // Blocks are dispatched in index order, so the predecessor of a waiting CTA is always resident or done.
  const int r = (int)(blockIdx.x % (unsigned)A.R), chunk = (int)(blockIdx.x / (unsigned)A.R);
  const int k0 = A.chunk_steps > 0 ? chunk * A.chunk_steps : 0;
  const int k1 = A.chunk_steps > 0 ? min(NT, k0 + A.chunk_steps) : NT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kThreads / 32;
  const uint32_t rglob = (uint32_t)(A.first_realization + r);

  extern __shared__ __align__(16) unsigned char smem[];
  const int npad = (n + 3) & ~3;
  real2* xbuf0 = reinterpret_cast<real2*>(smem);
  real2* xbuf1 = xbuf0 + npad;
  double* infx = reinterpret_cast<double*>(xbuf1 + npad);  // [2][L]  (cur, nxt)
  double* infy = infx + 2 * L;
  double* medx = infy + 2 * L;                             // [2][M]
  double* medy = medx + 2 * M;
  double* frac = medy + 2 * M;                             // [L]
  // bin sums of old positions: every bin is summed in two halves (agents a with (a >> 5) even / odd) so that
  // the 2(L+M) work items balance over the 4 warps; entry 2*bin + half
  double* bsx = frac + L;                                  // [2(L+M)]
  double* bsy = bsx + 2 * (L + M);
  int* bcnt = reinterpret_cast<int*>(bsy + 2 * (L + M));   // [2(L+M)]
  int* bss = bcnt + 2 * (L + M);                           // [2(L+M)] sum of states per bin
  int* fol = bss + 2 * (L + M);                            // [npad]
  float* finx = reinterpret_cast<float*>(fol + npad);      // [L] FP32 copies of the UPDATED influencers / fractions
  float* finy = finx + L;                                  //     for the guarded FP32 switch decision
  float* ffr = finy + L;
  signed char* st = reinterpret_cast<signed char*>(ffr + L);  // [npad]
  // SYM: positions once more, every 32-group stored twice in a row ([G][64]) so that the rotated partner
  // index lane +- s needs no wrap-around arithmetic, and the partial sums: slot d = 0..D: row sums of tile (g, g+d),
  // slot D + d, d = 1..D: column sums of tile (h-d, h); each [G*32] x (ws, ax, ay)
  const int G = (n + 31) >> 5, D = G >> 1, NG = G * 32;
  // xy2[g][i] = log2(e) * (x_a, x_b, y_a, y_b) of the agents a = g*32 + (i & 31), b = g*32 + ((i + 16) & 31), i < 64
  // wide CTAs: the tiles each warp evaluates (sched[warp][0] = count, then the tile indices), fixed at kernel start
  unsigned char* sched = reinterpret_cast<unsigned char*>(st + npad);
  float4* xy2 = reinterpret_cast<float4*>(smem + ((sched + (SYM && THREADS != 128 ? nwarps * kSchedMax : 0) - smem + 15) & ~(size_t)15));
  // SYM: class sums of the current state per warp, in fixed point: wsum[warp][bin] = (sum qx, sum qy, count, sum of states)
  // with q = round(position * qscale).  Integer sums are exact and associative: the class means do not depend on the
  // CTA width or on which warp holds which agent.
  int4* wsum = reinterpret_cast<int4*>(xy2 + (SYM ? G * kTab : 0));
  float* part = reinterpret_cast<float*>(wsum + (SYM ? nwarps * (L + M) : 0));

  // ---- load the initial condition, or the state the previous chunk of this realization left -------------------------
  if (chunk > 0) {
    const real* hand_x = static_cast<const real*>(A.hand_x) + (size_t)r * 2 * n;  // same layouts as x0 / follow0 / inf0 / media0
    const int32_t* hand_f = A.hand_f + (size_t)r * n;
    const double* hand_inf = A.hand_inf + (size_t)r * 2 * L;
    const double* hand_med = A.hand_med + (size_t)r * 2 * M;
    if (tid == 0) {
      unsigned done;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(A.chunk_done + r) : "memory");
        if (done < (unsigned)chunk) __nanosleep(200);
      } while (done < (unsigned)chunk);
    }
    __syncthreads();
    for (int j = tid; j < n; j += kThreads) {  // L2 loads: an earlier chunk of this realization may have run on this SM
      xbuf0[j] = mk2(__ldcg(hand_x + j), __ldcg(hand_x + n + j));
      fol[j] = __ldcg(hand_f + j);
      st[j] = A.state[(size_t)r * n + j];
    }
    if (tid < L) { infx[tid] = __ldcg(hand_inf + tid); infy[tid] = __ldcg(hand_inf + L + tid); }
    if (tid < M) { medx[tid] = __ldcg(hand_med + tid); medy[tid] = __ldcg(hand_med + M + tid); }
  } else {
    const real* x0 = static_cast<const real*>(A.x0) + (size_t)r * 2 * n;
    for (int j = tid; j < n; j += kThreads) {
      xbuf0[j] = mk2(x0[j], x0[n + j]);
      fol[j] = A.follow0[(size_t)r * n + j];
      st[j] = A.state[(size_t)r * n + j];
    }
    if (tid < L) { infx[tid] = A.inf0[(size_t)r * 2 * L + tid]; infy[tid] = A.inf0[(size_t)r * 2 * L + L + tid]; }
    if (tid < M) { medx[tid] = A.media0[(size_t)r * 2 * M + tid]; medy[tid] = A.media0[(size_t)r * 2 * M + M + tid]; }
  }
  __syncthreads();

  if constexpr (SYM && THREADS != 128) {
    // Tiles dealt to the warps once, by the longest-processing-time rule (off-diagonal tiles, 16 steps each, first; then
    // the diagonal ones, 9 steps), with warp 0 starting from the cost of the influencer / media update it also carries
    // (A.p2_cost steps): that update sits in front of warp 0's tiles, and every other warp waits for it at the barrier
    // behind the pair loop.  Every tile owns its partial-sum slots, so any assignment gives bit-identical results.
    if (tid == 0) {
      const int Gs = (n + 31) >> 5, Ds = Gs >> 1;
      const int NOs = (Gs & 1) ? Ds * Gs : (Ds > 0 ? (Ds - 1) * Gs + Ds : 0), NIs = NOs + Gs;
      int load[THREADS / 32];
      for (int w = 0; w < nwarps; w++) { load[w] = 0; sched[w * kSchedMax] = 0; }
      load[0] = A.p2_cost;
      for (int it = 0; it < NIs; it++) {
        int best = nwarps - 1;
        for (int w = nwarps - 2; w >= 0; w--)
          if (load[w] < load[best]) best = w;
        unsigned char* my = sched + best * kSchedMax;
        if (my[0] < kSchedMax - 1) { my[++my[0]] = (unsigned char)it; load[best] += it < NOs ? 16 : 9; }
        else { for (int w = 0; w < nwarps; w++) if (sched[w * kSchedMax] < kSchedMax - 1) { unsigned char* o = sched + w * kSchedMax; o[++o[0]] = (unsigned char)it; load[w] += 16; break; } }
      }
    }
    __syncthreads();
  }

  int cur = 0;        // which x buffer / inf slot holds the current state
  int next_save = 0;  // index into savesteps
  if (chunk > 0)      // history entry k0 was written by the previous chunk
    while (next_save < A.nsave && A.savesteps[next_save] <= k0) next_save++;
  const double sdt = sqrt(K.dt);
  const double noise_x = sdt * K.sigma;

  // K4 inside K1 (ensemble.jl:27-41): class densities of the snapshot straight from shared memory.  Agents are counting-
  // sorted by class c = 2 j + [state = +1] (stable: summation order = agent order, as in kde_kernel), staged 16 at a time as
  // separable factors ex[k][ix], ey[k][iy]; a thread accumulates a 12 x 3 tile in packed FP32x2 registers.  `other` is the
  // position buffer that is free between two steps, `part` the partial-sum area of the pair loop.
  auto kde_snapshot = [&](const real2* xc, real2* other, size_t s) {
    if constexpr (SYM) {
      const int ncls = 2 * L, Nx = A.kde_Nx, Ny = A.kde_Ny;
      int* off = bcnt;  // [ncls + 1] (the class-sum arrays of the ordered path are unused here)
      float* sax = reinterpret_cast<float*>(other);
      float* say = sax + n;
      float* ex = part;
      float* ey = ex + kKdeChunk * kKdeSpan;
      for (int c = tid; c <= ncls; c += kThreads) off[c] = 0;
      __syncthreads();
      for (int a = tid; a < n; a += kThreads) atomicAdd(&off[fol[a] * 2 + (st[a] == 1 ? 1 : 0) + 1], 1);
      __syncthreads();
      if (tid == 0)
        for (int c = 0; c < ncls; c++) off[c + 1] += off[c];
      __syncthreads();
      for (int c = warp; c < ncls; c += nwarps) {
        int pos = off[c];
        for (int a0 = 0; a0 < n; a0 += 32) {
          const int a = a0 + lane;
          const bool in = a < n && fol[a] * 2 + (st[a] == 1 ? 1 : 0) == c;
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) {
            const int q = pos + __popc(m & ((1u << lane) - 1u));
            const real2 pa = xc[a];
            sax[q] = (float)pa.x; say[q] = (float)pa.y;
          }
          pos += __popc(m);
        }
      }
      __syncthreads();
      const float inv2s2 = (float)(1.0 / (2.0 * A.kde_sigma * A.kde_sigma));
      const float scale = (float)(1.0 / (double)n / (2.0 * 3.14159265358979323846 * A.kde_sigma * A.kde_sigma));
      float* us = static_cast<float*>(A.kde_us) + s * ncls * Nx * Ny;
      for (int pass0 = 0; pass0 < kKdeSpan; pass0 += nwarps * kKdeRows) {
        const int row0 = pass0 + warp * kKdeRows;
        const bool active = row0 < kKdeSpan;
        for (int c = 0; c < ncls; c++) {
          float2 acc2[kKdeRows / 2][3];
#pragma unroll
          for (int b = 0; b < kKdeRows / 2; b++)
#pragma unroll
            for (int a = 0; a < 3; a++) acc2[b][a] = make_float2(0.f, 0.f);
          for (int k0 = off[c]; k0 < off[c + 1]; k0 += kKdeChunk) {
            const int kc = min(kKdeChunk, off[c + 1] - k0);
            __syncthreads();  // previous chunk fully consumed
            for (int e = tid; e < kc * kKdeSpan; e += kThreads) {
              const int k = e / kKdeSpan, i = e - k * kKdeSpan;
              ex[e] = i < Nx ? kde_gauss1d((float)(A.kde_gx0 + i * A.kde_dx) - sax[k0 + k], inv2s2) : 0.f;
              ey[e] = i < Ny ? kde_gauss1d((float)(A.kde_gy0 + i * A.kde_dy) - say[k0 + k], inv2s2) : 0.f;
            }
            __syncthreads();
            if (active)
              for (int k = 0; k < kc; k++) {
                float vx[3], vy[kKdeRows];
#pragma unroll
                for (int a = 0; a < 3; a++) vx[a] = ex[k * kKdeSpan + lane + 32 * a];
#pragma unroll
                for (int b = 0; b < kKdeRows; b += 4) {
                  const float4 v = *reinterpret_cast<const float4*>(ey + k * kKdeSpan + row0 + b);
                  vy[b] = v.x; vy[b + 1] = v.y; vy[b + 2] = v.z; vy[b + 3] = v.w;
                }
#pragma unroll
                for (int b = 0; b < kKdeRows; b += 2)
#pragma unroll
                  for (int a = 0; a < 3; a++) acc2[b / 2][a] = fma2(make_float2(vx[a], vx[a]), make_float2(vy[b], vy[b + 1]), acc2[b / 2][a]);
              }
          }
          if (active) {
            float* o = us + (size_t)c * Nx * Ny;
#pragma unroll
            for (int b = 0; b < kKdeRows; b++)
#pragma unroll
              for (int a = 0; a < 3; a++) {
                const int gx = lane + 32 * a, gy = row0 + b;
                if (gx < Nx && gy < Ny) o[(size_t)gy * Nx + gx] = ((b & 1) ? acc2[b / 2][a].y : acc2[b / 2][a].x) * scale;
              }
          }
        }
      }
      __syncthreads();  // the staging area is the pair loop's partial-sum area again
    }
  };

  auto snapshot = [&](int hist_index, const real2* xc, int slot) {
    // history entry `hist_index` (0 = initial state): abm.jl:161-164, 209-212
    if (A.infs) {
      double* o = A.infs + ((size_t)r * (NT + 1) + hist_index) * 2 * L;
      if (tid < L) { o[tid] = infx[slot * L + tid]; o[L + tid] = infy[slot * L + tid]; }
    }
    if (A.meds) {
      double* o = A.meds + ((size_t)r * (NT + 1) + hist_index) * 2 * M;
      if (tid < M) { o[tid] = medx[slot * M + tid]; o[M + tid] = medy[slot * M + tid]; }
    }
    if (next_save < A.nsave && A.savesteps[next_save] == hist_index) {
      size_t s = (size_t)r * A.nsave + next_save;
      if (A.xs) {
        real* o = static_cast<real*>(A.xs) + s * 2 * n;
        for (int j = tid; j < n; j += kThreads) { real2 p = xc[j]; o[j] = p.x; o[n + j] = p.y; }
      }
      if (A.follows)
        for (int j = tid; j < n; j += kThreads) A.follows[s * n + j] = fol[j];
      if (A.ys && tid < L) {  // Julia ys[2,L,sp,R] (ensemble.jl:39)
        A.ys[(s * L + tid) * 2 + 0] = infx[slot * L + tid];
        A.ys[(s * L + tid) * 2 + 1] = infy[slot * L + tid];
      }
      if (A.zs && tid < M) {  // Julia zs[2,2,sp,R] (ensemble.jl:41)
        A.zs[(s * M + tid) * 2 + 0] = medx[slot * M + tid];
        A.zs[(s * M + tid) * 2 + 1] = medy[slot * M + tid];
      }
      if (SYM && A.kde_us) kde_snapshot(xc, xc == xbuf0 ? xbuf1 : xbuf0, s);
      next_save++;
    }
  };

  // bin sums over the current followers: bins 0..L-1 by influencer, L..L+M-1 by medium
  auto bin_sums = [&](const real2* xc) {
    for (int item = warp; item < 2 * (L + M); item += nwarps) {
      const int bin = item >> 1;
      real sx = 0, sy = 0;  // `real` accumulation: Float64 in Float64 mode, FP32 (<= 8 terms per lane + a tree) in FP32 mode
      int c = 0, ss = 0;
      for (int a = lane + 32 * (item & 1); a < n; a += 64) {
        int s = st[a];
        bool in = bin < L ? (fol[a] == bin) : (s == (bin == L ? -1 : 1));
        if (in) { real2 p = xc[a]; sx += p.x; sy += p.y; c++; ss += s; }
      }
      sx = warp_sum(sx); sy = warp_sum(sy); c = warp_sum(c); ss = warp_sum(ss);
      if (lane == 0) { bsx[item] = (double)sx; bsy[item] = (double)sy; bcnt[item] = c; bss[item] = ss; }
    }
  };
  auto add_numbers = [&](int hist_index) {
    // plotting.jl:284-307: numbers[t,i,j] = #agents with medium i following influencer j
    if (A.numbers && tid < L * M) {
      int j = tid / M, i = tid % M;
      int c = bcnt[2 * j] + bcnt[2 * j + 1], ss = bss[2 * j] + bss[2 * j + 1];
      int v = i == 1 ? (c + ss) / 2 : (c - ss) / 2;
      atomicAdd(reinterpret_cast<unsigned long long*>(A.numbers) + ((size_t)hist_index * L + j) * M + i,
                (unsigned long long)v);
    }
  };

  // ---- SYM: the pair table and the class sums are kept up to date by the thread that moves an agent ------------------
  // table entry of agent j (see kTab): the scaled position goes to the four places the rotated reads expect
  auto table_put = [&](int j, float px, float py) {
    const float sx = px * kLog2e, sy = py * kLog2e;
    const int l = j & 31, l2 = (l + 16) & 31;
    float4* o = xy2 + (j >> 5) * 64;
    o[l].x = sx; o[l + 32].x = sx; o[l2].y = sx; o[l2 + 32].y = sx;
    o[l].z = sy; o[l + 32].z = sy; o[l2].w = sy; o[l2 + 32].w = sy;
  };
  // positions in fixed point: |q| <= 2^22, at most 128 agents per warp and pass: the warp sums fit 32 bits
  const float qscale = 4194304.0f / (float)fmax(fmax(fabs(K.lo0), fabs(K.hi0)), fmax(fabs(K.lo1), fabs(K.hi1)));
  int4 wacc = make_int4(0, 0, 0, 0);  // lane b: running sums of class b (b < L: followers of influencer b; then media)
  // one agent per lane (valid or not): every class summed over the warp with one REDUX per value; warp-uniform call
  auto class_add = [&](bool valid, float px, float py, int f, int s) {
    const int qx = __float2int_rn(px * qscale), qy = __float2int_rn(py * qscale);
    const int fe = valid ? f : -1;
    const unsigned plus = __ballot_sync(0xffffffffu, valid && s == 1);
    unsigned mine = 0u;  // lane b: the members of class b among this warp's agents
    for (int b = 0; b < L; b++) {
      const bool in = fe == b;
      const unsigned m = __ballot_sync(0xffffffffu, in);
      const int rx = __reduce_add_sync(0xffffffffu, in ? qx : 0), ry = __reduce_add_sync(0xffffffffu, in ? qy : 0);
      if (lane == b) { wacc.x += rx; wacc.y += ry; mine = m; }
    }
    for (int b = L; b < L + M; b++) {  // media classes by state: -1, +1 (abm.jl:188-191)
      const bool in = valid && s == (b == L ? -1 : 1);
      const unsigned m = __ballot_sync(0xffffffffu, in);
      const int rx = __reduce_add_sync(0xffffffffu, in ? qx : 0), ry = __reduce_add_sync(0xffffffffu, in ? qy : 0);
      if (lane == b) { wacc.x += rx; wacc.y += ry; mine = m; }
    }
    wacc.z += __popc(mine);
    wacc.w += 2 * __popc(mine & plus) - __popc(mine);
  };
  auto class_flush = [&]() {
    if (lane < L + M) wsum[warp * (L + M) + lane] = wacc;
    wacc = make_int4(0, 0, 0, 0);
  };
  // exact totals of class `lane` over the warps (warp 0 only; zero in the other lanes), then the follower numbers
  // (plotting.jl:284-307) from them
  long long tot_sx = 0, tot_sy = 0;
  int tot_c = 0, tot_ss = 0;
  auto class_totals = [&](int hist_index) {
    tot_sx = tot_sy = 0; tot_c = tot_ss = 0;
    if (lane < L + M)
      for (int w = 0; w < nwarps; w++) { const int4 v = wsum[w * (L + M) + lane]; tot_sx += v.x; tot_sy += v.y; tot_c += v.z; tot_ss += v.w; }
    if (A.numbers) {
      const int j = lane / M, i = lane % M;
      const int c = __shfl_sync(0xffffffffu, tot_c, j), ss = __shfl_sync(0xffffffffu, tot_ss, j);
      if (lane < L * M)
        atomicAdd(reinterpret_cast<unsigned long long*>(A.numbers) + ((size_t)hist_index * L + j) * M + i,
                  (unsigned long long)(i == 1 ? (c + ss) / 2 : (c - ss) / 2));
    }
  };
  if constexpr (SYM) {
    static_assert(sizeof(real) == 4, "the symmetric path is FP32");
    for (int j = tid; j < NG; j += kThreads) {
      const real2 p = j < n ? xbuf0[j] : mk2((real)kPadPos, (real)kPadPos);  // padding agents: written once
      table_put(j, (float)p.x, (float)p.y);
    }
    for (int jb = 0; jb < n; jb += kThreads) {
      const int j = jb + tid;
      const bool valid = j < n;
      const real2 p = xbuf0[valid ? j : 0];
      class_add(valid, (float)p.x, (float)p.y, fol[valid ? j : 0], st[valid ? j : 0]);
    }
    class_flush();
    __syncthreads();
  }

  if (chunk == 0) snapshot(0, xbuf0, 0);

#ifdef SMM_DENSE_STAMPS
  long long st_loop = 0, st_w1 = 0, st_fin = 0, st_w2 = 0;
  const long long st_c0 = clock64();
  unsigned long long st_g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_g0));
#endif
  for (int k = k0; k < k1; k++) {
#ifdef SMM_DENSE_STAMPS
    const long long t_a = clock64();
#endif
    const real2* xc = cur ? xbuf1 : xbuf0;
    real2* xn = cur ? xbuf0 : xbuf1;
    const int nxt = cur ^ 1;

    // ---- P1: S_i, cnt_i, state sums of the OLD positions / followers (abm.jl:96-100,179-180,188-191)
    if constexpr (SYM) {
      if (warp == 0) class_totals(k);  // the class sums of the current state were left by the previous step (or the prologue)
    } else {
      bin_sums(xc);
      __syncthreads();
      add_numbers(k);
    }

    // ---- P2: influencer / media update from OLD x, OLD followers (abm.jl:175-198), reflect (:202-203)
    if (tid < L + M) {
      const bool is_inf = tid < L;
      const int e = is_inf ? tid : tid - L;
      double px = is_inf ? infx[cur * L + e] : medx[cur * M + e];
      double py = is_inf ? infy[cur * L + e] : medy[cur * M + e];
      const double fr = is_inf ? K.frictionI : K.frictionM;
      const double sg = is_inf ? K.sigmahat : K.sigmatilde;
      double z0 = 0.0, z1 = 0.0;
      if (sg != 0.0) {
        if (A.noise_mode == SMM_NOISE_INJECTED) {
          const double* zp = is_inf ? A.xi_inf : A.xi_med;
          if (zp) {
            size_t o = (((size_t)r * NT + k) * (is_inf ? L : M) + e) * 2;
            z0 = zp[o]; z1 = zp[o + 1];
          }
        } else {
          float f0, f1;
          entity_draws(A.seed, (uint32_t)tid, (uint32_t)k, rglob, f0, f1);
          z0 = f0; z1 = f1;
        }
      }
      const double nz = 1.0 / fr * sdt * sg;
      int cnt, ssum;
      double sumx, sumy;
      if constexpr (SYM) {
        cnt = tot_c; ssum = tot_ss; sumx = (double)tot_sx / (double)qscale; sumy = (double)tot_sy / (double)qscale;
      } else {
        cnt = bcnt[2 * tid] + bcnt[2 * tid + 1]; ssum = bss[2 * tid] + bss[2 * tid + 1];
        sumx = bsx[2 * tid] + bsx[2 * tid + 1]; sumy = bsy[2 * tid] + bsy[2 * tid + 1];
      }
      if (cnt > 0) {
        double mcx = sumx / (double)cnt, mcy = sumy / (double)cnt;
        px = px + K.dt / fr * (mcx - px) + nz * z0;
        py = py + K.dt / fr * (mcy - py) + nz * z1;
      } else {
        px = px + nz * z0;
        py = py + nz * z1;
      }
      px = reflect1(px, K.lo0, K.hi0);
      py = reflect1(py, K.lo1, K.hi1);
      if (is_inf) {
        infx[nxt * L + e] = px; infy[nxt * L + e] = py;
        const double fr_e = (double)ssum / (double)cnt;  // abm.jl:99; 0/0 = NaN (Q5)
        frac[e] = fr_e;
        finx[e] = (float)px; finy[e] = (float)py; ffr[e] = (float)fr_e;
      } else {
        medx[nxt * M + e] = px; medy[nxt * M + e] = py;
      }
    }
    // The symmetric pair loop reads nothing P2 writes (its consumers are the per-agent updates behind the next barrier):
    // no barrier here, the six P2 threads' latency overlaps the other warps' tiles.
    if (!SYM) __syncthreads();

    // ---- P3: per agent: attraction over all agents, influence, Euler-Maruyama, switch.
    // Two agents per thread (j0, j0 + kThreads): each broadcast partner load feeds two pair
    // evaluations, and 128-thread CTAs let every realization of a 1 024-member ensemble be
    // resident at once (7 CTAs per SM) instead of two waves of 256-thread CTAs.
    real2 moved = mk2((real)0, (real)0);  // finish(): the agent's new position (as stored) and follower
    int fmoved = 0;
    auto finish = [&](int j, real2 xj, real ax, real ay, real ws) {
      const int sj = st[j], fj = fol[j];
      const double x = (double)xj.x, y = (double)xj.y;
      const int mrow = sj == 1 ? 1 : 0;  // abm.jl:74-78
      double fx = K.a * (double)(ax / ws) + K.c * (medx[cur * M + mrow] - x) + K.b * (infx[cur * L + fj] - x);
      double fy = K.a * (double)(ay / ws) + K.c * (medy[cur * M + mrow] - y) + K.b * (infy[cur * L + fj] - y);
      double xi0, xi1, u1, u2;
      if (A.noise_mode == SMM_NOISE_INJECTED) {
        size_t o = ((size_t)r * NT + k);
        const real* xp = static_cast<const real*>(A.xi) + o * 2 * n;
        xi0 = (double)xp[j]; xi1 = (double)xp[n + j];
        u1 = A.u1[o * n + j]; u2 = A.u2[o * n + j];
      } else {
        float f0, f1, g1, g2;
        agent_draws(A.seed, (uint32_t)j, (uint32_t)k, rglob, f0, f1, g1, g2);
        xi0 = f0; xi1 = f1; u1 = g1; u2 = g2;
      }
      double xn0 = x + K.dt * fx + noise_x * xi0;  // abm.jl:173
      double xn1 = y + K.dt * fy + noise_x * xi1;
      xn0 = reflect1(xn0, K.lo0, K.hi0);           // abm.jl:201
      xn1 = reflect1(xn1, K.lo1, K.hi1);
      // abm.jl:207: OLD x, OLD follower fractions, UPDATED influencers (Q1)
      // Philox uniforms are 24-bit (exact in FP32): decide in FP32 with a guard band, Float64 inside it
      int fnew = fj;
      bool decided = false;
      if (sizeof(real) == 4 && A.noise_mode != SMM_NOISE_INJECTED)
        decided = switch_decision_f32_guarded((float)xj.x, (float)xj.y, sj, fj, finx, finy, ffr, L, (float)K.eta, (float)K.dt,
                                              (float)u1, (float)u2, 2e-5f, &fnew);
      if (!decided) fnew = switch_decision_f64(x, y, sj, fj, infx + nxt * L, infy + nxt * L, frac, L, K.eta, K.dt, u1, u2);
      moved = mk2((real)xn0, (real)xn1);
      fmoved = fnew;
      xn[j] = moved;
      fol[j] = fnew;  // only this thread reads fol[j] during P3
    };
    // SYM: partial sums of agent j, then finish(); the thread that moved the agent also writes its table entry for the next
    // step (the pair loop is behind a barrier) and adds it to the class sums of the next step.  SLOTS = number of slot
    // groups of the CTA form (row sums only / row and column sums).
    auto finish_all = [&](auto total_of) {
      for (int jb = 0; jb < n; jb += kThreads) {
        const int j = jb + tid;
        const bool valid = j < n;
        if (valid) {
          real ws, ax, ay;
          total_of(j, ws, ax, ay);
          finish(j, xc[j], ax, ay, ws);
          table_put(j, (float)moved.x, (float)moved.y);
        }
        class_add(valid, (float)moved.x, (float)moved.y, fmoved, st[valid ? j : 0]);
      }
      class_flush();
    };
    if constexpr (SYM) {
      // Attraction with every unordered pair evaluated once.  Agents are split into G groups of 32; a warp
      // owns row groups g = warp, warp + 4, ..: lane l is agent g*32 + l.  For the group offsets d = 1..D the
      // tile (g, h = g + d mod G) pairs row agent l with column agent (l + s) & 31 at rotation step s: the
      // weight goes into the row agent's registers and, through one shuffle, to the lane that owns the column
      // agent (which re-reads the row position from shared memory).  Offsets 1..D cover every unordered group
      // pair exactly once (for even G the offset G/2 only from the lower half); the diagonal tile (g, g) is
      // evaluated in ordered form and contains the self pair (abm.jl:42).  Every partial sum has its own
      // shared-memory slot and is added in a fixed order: results are deterministic.
      if constexpr (THREADS == 128) {
      // full-GPU configuration: a warp owns whole row groups and keeps their running totals in registers across the
      // offsets.  The totals are added in the order of the tile-dealing form below (diagonal tile, then the row sums
      // of d = 1..D, every tile folded on its own first), so that a realization's result does not depend on the CTA
      // width, i.e. on the size of the ensemble it is part of.
      for (int g = warp; g < G; g += nwarps) {
        const float4* rowp = xy2 + g * kTab + lane;
        const float rx = rowp[0].x, ry = rowp[0].z;
        const float2 nrx = make_float2(-rx, -rx), nry = make_float2(-ry, -ry);
        Sums2 Rs;
        diag_tile(rowp, lane, nrx, nry, Rs);
        float totw = fold_w(Rs), totx = fold_x(Rs), toty = fold_y(Rs);
        for (int d = 1; d <= D; d++) {
          if (2 * d == G && g >= D) break;
          int h = g + d;
          if (h >= G) h -= G;
          const float4* colp = xy2 + h * kTab + lane;
          const float cx = colp[0].x, cy = colp[0].z;
          Sums2 Cs;
          zero(Rs); zero(Cs);
          sym_steps<0, 15>(rowp, colp, lane, nrx, nry, make_float2(-cx, -cx), make_float2(-cy, -cy), Rs, Cs);
          float* o = part + (size_t)d * 3 * NG + h * 32 + lane;
          o[0] = fold_w(Cs); o[NG] = fold_x(Cs); o[2 * NG] = fold_y(Cs);
          totw = __fadd_rn(totw, fold_w(Rs)); totx = __fadd_rn(totx, fold_x(Rs)); toty = __fadd_rn(toty, fold_y(Rs));
        }
        float* o = part + g * 32 + lane;
        o[0] = totw; o[NG] = totx; o[2 * NG] = toty;
      }
#ifdef SMM_DENSE_STAMPS
      const long long t_b = clock64();
      __syncthreads();
      const long long t_c = clock64();
#else
      __syncthreads();
#endif
      finish_all([&](int j, real& ws, real& ax, real& ay) {
        ws = part[j]; ax = part[NG + j]; ay = part[2 * NG + j];
        const int hg = j >> 5;
        for (int d = 1; d <= D; d++) {
          if (2 * d == G && hg < D) break;  // that slot is only written for the upper half of the groups
          const real* o = part + (size_t)d * 3 * NG + j;
          ws += o[0]; ax += o[NG]; ay += o[2 * NG];
        }
      });
#ifdef SMM_DENSE_STAMPS
      const long long t_d = clock64();
      __syncthreads();
      const long long t_e = clock64();
      st_loop += t_b - t_a; st_w1 += t_c - t_b; st_fin += t_d - t_c; st_w2 += t_e - t_d;
#endif
      } else {
      // Work items = tiles: NO off-diagonal tiles (d = 1..D; for even G the offset G/2 only from the lower half of the
      // groups) and G diagonal ones.  Off-diagonal tiles cost ~1.7x a diagonal one, so they are dealt first and in
      // snake order over the warps (longest-processing-time rule): 36 tiles at n = 250 balance to within 6 % over 16
      // warps.  Every tile owns its partial-sum slots, so any assignment gives bit-identical results.
      const int NO = (G & 1) ? D * G : (D > 0 ? (D - 1) * G + D : 0);
      const unsigned char* mine = sched + warp * kSchedMax;
      const int nmine = mine[0];
      for (int si = 1; si <= nmine; si++) {
        const int it = mine[si];
        const int d = it < NO ? it / G + 1 : 0;
        const int g = it < NO ? it - (d - 1) * G : it - NO;
        const float4* rowp = xy2 + g * kTab + lane;
        const float rx = rowp[0].x, ry = rowp[0].z;
        const float2 nrx = make_float2(-rx, -rx), nry = make_float2(-ry, -ry);
        Sums2 Rs;
        if (d == 0) {
          diag_tile(rowp, lane, nrx, nry, Rs);
        } else {
          int h = g + d;
          if (h >= G) h -= G;
          const float4* colp = xy2 + h * kTab + lane;
          const float cx = colp[0].x, cy = colp[0].z;
          Sums2 Cs;
          zero(Rs); zero(Cs);
          sym_steps<0, 15>(rowp, colp, lane, nrx, nry, make_float2(-cx, -cx), make_float2(-cy, -cy), Rs, Cs);
          float* o = part + (size_t)(D + d) * 3 * NG + h * 32 + lane;
          o[0] = fold_w(Cs); o[NG] = fold_x(Cs); o[2 * NG] = fold_y(Cs);
        }
        float* o = part + (size_t)d * 3 * NG + g * 32 + lane;
        o[0] = fold_w(Rs); o[NG] = fold_x(Rs); o[2 * NG] = fold_y(Rs);
      }
      __syncthreads();
      finish_all([&](int j, real& ws, real& ax, real& ay) {
        ws = part[j]; ax = part[NG + j]; ay = part[2 * NG + j];
        const int hg = j >> 5;
        for (int d = 1; d <= D; d++) {  // fixed order: row sums d = 1..D, then column sums d = 1..D
          if (2 * d == G && hg >= D) break;  // tile (hg, hg + G/2) belongs to the lower half of the groups
          const real* o = part + (size_t)d * 3 * NG + j;
          ws += o[0]; ax += o[NG]; ay += o[2 * NG];
        }
        for (int d = 1; d <= D; d++) {
          if (2 * d == G && hg < D) break;   // column sums of offset G/2 exist only for the upper half
          const real* o = part + (size_t)(D + d) * 3 * NG + j;
          ws += o[0]; ax += o[NG]; ay += o[2 * NG];
        }
      });
      }
    } else
    for (int j0 = tid; j0 < n; j0 += 2 * kThreads) {
      const int j1 = j0 + kThreads;
      const bool has1 = j1 < n;
      const real2 xa = xc[j0], xb = xc[has1 ? j1 : j0];
      real ax0 = 0, ay0 = 0, ws0 = 0, ax1 = 0, ay1 = 0, ws1 = 0;
      if (A.rowptr == nullptr) {
#pragma unroll 4
        for (int i = 0; i < n; i++) {  // abm.jl:57-62; every lane reads the same address: smem broadcast
          const real2 p = xc[i];
          const real dx0 = p.x - xa.x, dy0 = p.y - xa.y, dx1 = p.x - xb.x, dy1 = p.y - xb.y;
          const real w0 = w_of_r2(dx0 * dx0 + dy0 * dy0), w1 = w_of_r2(dx1 * dx1 + dy1 * dy1);
          ws0 += w0; ax0 += w0 * dx0; ay0 += w0 * dy0;
          ws1 += w1; ax1 += w1 * dx1; ay1 += w1 * dy1;
        }
      } else {  // sparse follower network shared by the realizations (row j of IndNet, abm.jl:51)
        for (int e = A.rowptr[j0]; e < A.rowptr[j0 + 1]; e++) {
          const real2 p = xc[A.colidx[e]];
          const real dx = p.x - xa.x, dy = p.y - xa.y, w = w_of_r2(dx * dx + dy * dy);
          ws0 += w; ax0 += w * dx; ay0 += w * dy;
        }
        if (has1)
          for (int e = A.rowptr[j1]; e < A.rowptr[j1 + 1]; e++) {
            const real2 p = xc[A.colidx[e]];
            const real dx = p.x - xb.x, dy = p.y - xb.y, w = w_of_r2(dx * dx + dy * dy);
            ws1 += w; ax1 += w * dx; ay1 += w * dy;
          }
        if (ws0 == (real)0) ws0 = (real)1;  // no neighbours: zero force (abm.jl:52-53)
        if (ws1 == (real)0) ws1 = (real)1;
      }
      finish(j0, xa, ax0, ay0, ws0);
      if (has1) finish(j1, xb, ax1, ay1, ws1);
    }
    __syncthreads();
    cur = nxt;
    snapshot(k + 1, xn, cur);
  }
#ifdef SMM_DENSE_STAMPS
  if (THREADS == 128 && SYM && tid == 0) {
    unsigned long long st_g1;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_g1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("cta %d chunk %d sm %u start %llu end %llu loop %lld update %lld\n", r, chunk, smid, st_g0, st_g1, st_loop / max(1, k1 - k0), st_fin / max(1, k1 - k0));
  }
#endif
  if (k1 < NT) {  // hand the state to the next chunk of this realization
    const real2* xc = cur ? xbuf1 : xbuf0;
    real* hand_x = static_cast<real*>(A.hand_x) + (size_t)r * 2 * n;
    int32_t* hand_f = A.hand_f + (size_t)r * n;
    double* hand_inf = A.hand_inf + (size_t)r * 2 * L;
    double* hand_med = A.hand_med + (size_t)r * 2 * M;
    for (int j = tid; j < n; j += kThreads) { const real2 p = xc[j]; hand_x[j] = p.x; hand_x[n + j] = p.y; hand_f[j] = fol[j]; }
    if (tid < L) { hand_inf[tid] = infx[cur * L + tid]; hand_inf[L + tid] = infy[cur * L + tid]; }
    if (tid < M) { hand_med[tid] = medx[cur * M + tid]; hand_med[M + tid] = medy[cur * M + tid]; }
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.chunk_done + r), "r"((unsigned)(chunk + 1)) : "memory");
    return;
  }
  if constexpr (SYM) {
    if (warp == 0) class_totals(NT);
  } else if (A.numbers) {
    bin_sums(cur ? xbuf1 : xbuf0);
    __syncthreads();
    add_numbers(NT);
  }
}

// ABMinit on the device (abm.jl:8-45).  One CTA per realization.
template <typename real>
__global__ void __launch_bounds__(kThreads) abm_init_kernel(AbmInitArgs A) {
  const int n = A.k.n, L = A.k.L, M = A.k.M;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kThreads / 32;
  const uint32_t rglob = (uint32_t)(A.first_realization + r);
  real* x0 = static_cast<real*>(A.x0) + (size_t)r * 2 * n;
  const int g = A.cells;  // L = g*g
  for (int j = tid; j < n; j += kThreads) {
    uint4 b = philox4x32_10(make_uint4((uint32_t)j, 0u, SMM_STREAM_INIT, rglob),
                            make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)));
    // abm.jl:16: rand(n,2).*(domain[1,2]-domain[1,1]) .+ domain[1,1]  (row 1 of the domain for both coordinates)
    double w = A.k.hi0 - A.k.lo0;
    real x = (real)((double)u24(b.x) * w + A.k.lo0);
    real y = (real)((double)u24(b.y) * w + A.k.lo0);
    x0[j] = x; x0[n + j] = y;
    A.state[(size_t)r * n + j] = (u24(b.z) > 0.5f) ? 1 : -1;  // abm.jl:39
    int f;
    if (g == 2) {  // abm.jl:21-24: 1:(+,+) 2:(<=0,+) 3:(+,<=0) 4:(<=0,<=0)
      f = (y > 0) ? ((x > 0) ? 0 : 1) : ((x > 0) ? 2 : 3);
    } else {       // g x g partition of the domain (C5: L = 64), row-major in (y cell, x cell)
      int cx = min(g - 1, max(0, (int)floor(((double)x - A.k.lo0) / w * g)));
      int cy = min(g - 1, max(0, (int)floor(((double)y - A.k.lo0) / w * g)));
      f = cy * g + cx;
    }
    A.follow0[(size_t)r * n + j] = f;
  }
  __syncthreads();
  // abm.jl:33-36: influencer = mean of its followers (NaN if none)
  for (int i = warp; i < L; i += nwarps) {
    double sx = 0, sy = 0; int c = 0;
    for (int a = lane; a < n; a += 32)
      if (A.follow0[(size_t)r * n + a] == i) { sx += (double)x0[a]; sy += (double)x0[n + a]; c++; }
    sx = warp_sum(sx); sy = warp_sum(sy); c = warp_sum(c);
    if (lane == 0) {
      A.inf0[(size_t)r * 2 * L + i] = sx / (double)c;
      A.inf0[(size_t)r * 2 * L + L + i] = sy / (double)c;
    }
  }
  if (tid < M) {  // abm.jl:18: media = [-1 -1; 1 1]
    A.media0[(size_t)r * 2 * M + tid] = tid == 0 ? -1.0 : 1.0;
    A.media0[(size_t)r * 2 * M + M + tid] = tid == 0 ? -1.0 : 1.0;
  }
}

__global__ void abm_draws_kernel(int n, uint64_t seed, uint32_t realization, int NT, double* xi, double* u1, double* u2) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * NT) return;
  int k = (int)(t / n), j = (int)(t % n);
  float a, b, c, d;
  agent_draws(seed, (uint32_t)j, (uint32_t)k, realization, a, b, c, d);
  xi[((size_t)k * 2 + 0) * n + j] = a;
  xi[((size_t)k * 2 + 1) * n + j] = b;
  u1[t] = c;
  u2[t] = d;
}

size_t dense_smem_bytes(int n, int L, int M, size_t real_size, bool sym, int threads, bool kde = false) {
  int npad = (n + 3) & ~3;
  size_t b = 2 * (size_t)npad * 2 * real_size + sizeof(double) * (4 * L + 4 * M + L + 4 * (L + M)) +
             sizeof(int) * (4 * (L + M) + npad) + sizeof(float) * 3 * L + npad + 16;
  if (sym) {
    const size_t G = (n + 31) / 32, D = G / 2;
    // 128-thread CTAs keep row sums in registers (slots 0..D); wider CTAs deal single tiles (row + column slots).
    // The difference decides whether 7 CTAs (all 1 024 realizations of C2) fit one SM's shared memory.
    const size_t part = (1 + (threads == 128 ? 1 : 2) * D) * 3 * G * 32 * 4, stage = kde ? 2 * kKdeChunk * kKdeSpan * 4 : 0;
    if (threads != 128) b += (size_t)(threads / 32) * kSchedMax;
    b += 16 + G * kTab * 16 + (size_t)(threads / 32) * (L + M) * 16 + std::max(part, stage);
  }
  return b;
}

}  // namespace

template <typename real, bool SYM, int THREADS>
static int launch_dense_t(smm_ctx* ctx, const AbmDenseArgs& A, size_t sm) {
  SMM_CUDA(ctx, cudaFuncSetAttribute(abm_dense_kernel<real, SYM, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int nchunks = A.chunk_steps > 0 ? (A.NT + A.chunk_steps - 1) / A.chunk_steps : 1;
  abm_dense_kernel<real, SYM, THREADS><<<(unsigned)A.R * (unsigned)std::max(1, nchunks), THREADS, sm, ctx->stream>>>(A);
  return SMM_OK;
}
template <typename real, bool SYM>
static int launch_dense_w(smm_ctx* ctx, const AbmDenseArgs& A, size_t sm, int threads) {
  if (threads == 512) return launch_dense_t<real, SYM, 512>(ctx, A, sm);
  if (threads == 256) return launch_dense_t<real, SYM, 256>(ctx, A, sm);
  return launch_dense_t<real, SYM, 128>(ctx, A, sm);
}

static bool dense_takes_sym_path(const AbmDenseArgs& A, int precision) {
  // FP32 all-to-all populations of up to 512 agents take the symmetric-pair path (SMM_ABM_SYM=0: ordered pairs)
  const char* ev = getenv("SMM_ABM_SYM");
  return precision == SMM_PREC_F32 && A.rowptr == nullptr && (A.k.n + 31) / 32 <= kSymMaxGroups && A.k.L + A.k.M <= 32 && A.k.L * A.k.M <= 32 &&
         !(ev && ev[0] == '0');
}

bool smm_abm_dense_fuses_kde(const smm_ctx* ctx, const AbmDenseArgs& A, int precision, int Nx, int Ny) {
  // Needs the symmetric path, a grid of one 96 x 96 pass and class offsets that fit the small-array area.  Opt-in
  // (SMM_ABM_FUSED_KDE=1): same bits as kde_kernel, but measured slower at 1 024 realizations -- the step kernel grows by
  // 0.90 ms (128-thread CTAs need two passes over the grid and stage the separable factors twice) where kde_kernel takes
  // 0.71 ms: 6.24 against 6.06 ms per pass.  The step kernel does not hide added work.
  const char* ev = getenv("SMM_ABM_FUSED_KDE");
  return dense_takes_sym_path(A, precision) && Nx <= kKdeSpan && Ny <= kKdeSpan && A.k.M >= 1 && ev && ev[0] == '1';
}

int smm_launch_abm_dense(smm_ctx* ctx, const AbmDenseArgs& A, int precision) {
  const bool sym = dense_takes_sym_path(A, precision);
  SMM_REQUIRE(ctx, A.kde_us == nullptr || (sym && A.kde_Nx <= kKdeSpan && A.kde_Ny <= kKdeSpan), "dense ABM kernel: fused KDE needs the symmetric path");
  SMM_REQUIRE(ctx, A.k.L + A.k.M <= kThreads && A.k.L * A.k.M <= kThreads, "dense ABM kernel: L+M and L*M must be <= %d", kThreads);
  // CTA width from the ensemble size (never wider than the tile count of the population can feed: n = 250 has 8 row
  // groups, 36 tiles).  With step chunks (below) the realizations need not all be resident at once, they take turns on
  // the CTA slots chunk by chunk, so the width is chosen for throughput.  Measured per pass (200 steps, n = 250, chunks of
  // 20 steps; ms): R = 170: 512 threads 1.64, 256 threads 1.98; R = 200: 1.91 / 2.08; R = 256: 256 threads 2.06 (one
  // chunk); R = 370: 256 threads 2.44, 128 threads 3.14; R = 512: 3.06 / 3.41; R = 600: 3.58 / 3.63; R = 700: 4.18 / 3.99;
  // R = 1 024: 6.07 / 5.20 (profiles/r2c_dense_chunk_sweep.txt).
  int threads = 128;
  {
    const int groups = (A.k.n + 31) / 32, sms = ctx->sm_count;
    if (2 * A.R <= 3 * sms && groups >= 8) threads = 512;
    else if (A.R <= 4 * sms && groups >= 4) threads = 256;
    if (const char* tv = getenv("SMM_ABM_THREADS")) { const int t = atoi(tv); if (t == 128 || t == 256 || t == 512) threads = t; }
  }
  // Step chunks (see the kernel) where the ensemble has more realizations than one round of CTA slots can balance: more
  // than two per SM (one for 512-thread CTAs, which have one slot per SM).  Below that every realization has a slot of its
  // own for the whole run and a chunk boundary only costs (R = 256: 2.06 -> 2.20 ms).  Measured at 1 024 realizations x 200
  // steps: one chunk 5.68 ms, chunks of 50 / 25 / 20 / 10 / 5 steps 5.31 / 5.22 / 5.20 / 5.19 / 5.22 ms (the prologue --
  // pair table and class sums -- and the hand-over cost about one step per chunk); R = 330: 2.77 -> 2.32 ms; R = 600:
  // 4.28 -> 3.62 ms.  SMM_ABM_CHUNK_STEPS overrides (0 = one chunk).
  AbmDenseArgs B = A;
  B.chunk_steps = 0;
  if (A.NT >= 40 && A.R > (threads == 512 ? 1 : 2) * ctx->sm_count) B.chunk_steps = 20;
  // measured (ms per pass, R = 128 / 296 realizations): snake order without head start 1.345 / 2.069; LPT with head start
  // 0 / 10 / 20 / 30 steps: 1.281 / 1.240 / 1.248 / 1.251 and 2.063 / 2.025 / 2.029 / 2.037
  B.p2_cost = 10;
  if (const char* pv = getenv("SMM_ABM_P2_COST")) B.p2_cost = std::max(0, atoi(pv));
  if (const char* cv = getenv("SMM_ABM_CHUNK_STEPS")) B.chunk_steps = std::max(0, atoi(cv));
  if (B.chunk_steps >= A.NT) B.chunk_steps = 0;
  if (B.chunk_steps > 0) {
    const size_t R = (size_t)A.R, n = (size_t)A.k.n, L = (size_t)A.k.L, M = (size_t)A.k.M, rs = precision == SMM_PREC_F64 ? 8 : 4;
    void* h = nullptr;
    SMM_TRY(smm_scratch(ctx, S_HANDOFF, R * (2 * L + 2 * M) * 8 + R * 2 * n * rs + R * n * 4 + R * 4, &h));
    B.hand_inf = (double*)h;  // widest element type first: every array is aligned for its type
    B.hand_med = B.hand_inf + R * 2 * L;
    B.hand_x = (void*)(B.hand_med + R * 2 * M);
    B.hand_f = (int32_t*)((char*)B.hand_x + R * 2 * n * rs);
    B.chunk_done = (unsigned*)(B.hand_f + R * n);
    SMM_CUDA(ctx, cudaMemsetAsync(B.chunk_done, 0, R * 4, ctx->stream));
  }
  size_t sm = dense_smem_bytes(A.k.n, A.k.L, A.k.M, precision == SMM_PREC_F64 ? 8 : 4, sym, threads, A.kde_us != nullptr);
  SMM_REQUIRE(ctx, sm <= 227 * 1024, "dense ABM kernel: n=%d does not fit shared memory (%zu B); pass a CSR network", A.k.n, sm);
  int rc;
  if (precision == SMM_PREC_F64) rc = launch_dense_w<double, false>(ctx, B, sm, threads);
  else if (sym) rc = launch_dense_w<float, true>(ctx, B, sm, threads);
  else rc = launch_dense_w<float, false>(ctx, B, sm, threads);
  SMM_TRY(rc);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_abm_init(smm_ctx* ctx, const AbmInitArgs& A, int R, int precision) {
  if (precision == SMM_PREC_F64) abm_init_kernel<double><<<R, kThreads, 0, ctx->stream>>>(A);
  else abm_init_kernel<float><<<R, kThreads, 0, ctx->stream>>>(A);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_launch_abm_draws(smm_ctx* ctx, int n, uint64_t seed, uint32_t realization, int NT, double* xi, double* u1, double* u2) {
  size_t tot = (size_t)n * NT;
  abm_draws_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(n, seed, realization, NT, xi, u1, u2);
  ctx->launches++;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}
