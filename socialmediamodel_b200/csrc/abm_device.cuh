// abm_device.cuh -- device functions shared by the dense (block-per-realization) and the
// CSR (population) ABM kernels.
#pragma once
#include "smm_common.cuh"

// changeinfluencer for one agent (abm.jl:103-133) in Float64 with the reference's operation
// order (no FMA contraction), so that under injected uniforms the sampled influencer is the
// reference's.  x,y: OLD position; infx/infy: UPDATED, reflected influencers (Q1 aliasing,
// abm.jl:175,207); frac[i] from the OLD follower network (abm.jl:96-100; NaN if influencer i
// has no follower, which makes lam NaN and the comparison false: Q5).
__device__ __forceinline__ int switch_decision_f64(double x, double y, int st, int follow, const double* infx,
                                                   const double* infy, const double* frac, int L, double eta,
                                                   double dt, double u1, double u2) {
  double lam = 0.0;
  for (int i = 0; i < L; i++) {
    double d0 = x - infx[i], d1 = y - infy[i];
    double dist = exp(-sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1))));
    double g2 = __dmul_rn((double)st, frac[i]);
    if (g2 < SMM_THETA) g2 = SMM_THETA;
    lam = __dadd_rn(lam, __dmul_rn(__dmul_rn(eta, dist), g2));
  }
  double thr = 1.0 - exp(-__dmul_rn(lam, dt));
  if (!(u1 < thr)) return follow;  // abm.jl:124 (false for NaN)
  double cs = 0.0;
  int k = L - 1;  // Q7: clamp
  for (int i = 0; i < L; i++) {
    double d0 = x - infx[i], d1 = y - infy[i];
    double dist = exp(-sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1))));
    double g2 = __dmul_rn((double)st, frac[i]);
    if (g2 < SMM_THETA) g2 = SMM_THETA;
    double att = __dmul_rn(__dmul_rn(eta, dist), g2);
    cs = __dadd_rn(cs, att / lam);  // abm.jl:125,127
    if (!(cs < u2)) { k = i; break; }
  }
  return k;
}

// FP32 evaluation of the same decision with a guard band: returns true and sets *out when
// every comparison is farther from its threshold than `band` (relative to the compared
// quantities), so the FP64 routine would decide identically; otherwise returns false and the
// caller falls back to switch_decision_f64.  Used where L is large (C5: L = 64).
__device__ __forceinline__ bool switch_decision_f32_guarded(float x, float y, int st, int follow, const float* infx,
                                                            const float* infy, const float* frac, int L, float eta,
                                                            float dt, float u1, float u2, float band, int* out) {
  float lam = 0.f;
  for (int i = 0; i < L; i++) {
    float d0 = x - infx[i], d1 = y - infy[i];
    float g2 = (float)st * frac[i];
    if (g2 < 0.1f) g2 = 0.1f;
    lam += eta * w_of_r2(fmaf(d0, d0, d1 * d1)) * g2;
  }
  if (!(lam == lam)) { *out = follow; return true; }  // NaN lock (Q5)
  float e = lam * dt;
  // 1-exp(-e) in FP32; e is O(0.01..1) so no cancellation problem beyond ~1e-7 absolute
  float thr = -expm1f(-e);
  float gap = u1 - thr;
  if (fabsf(gap) <= band * fmaxf(thr, 1e-3f)) return false;
  if (gap >= 0.f) { *out = follow; return true; }
  float cs = 0.f;
  float inv = 1.0f / lam;
  for (int i = 0; i < L - 1; i++) {
    float d0 = x - infx[i], d1 = y - infy[i];
    float g2 = (float)st * frac[i];
    if (g2 < 0.1f) g2 = 0.1f;
    cs += eta * w_of_r2(fmaf(d0, d0, d1 * d1)) * g2 * inv;
    float g = cs - u2;
    if (fabsf(g) <= band) return false;
    if (g >= 0.f) { *out = i; return true; }
  }
  *out = L - 1;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative switch decision for many influencers (population kernel, L up to 256).
//   phase 1 (per lane, all lanes active): lam = sum_i att_i in FP32, thr = 1-exp(-lam dt);
//           |u1-thr| outside the guard band decides "stay"/"switch"; inside -> FP64.
//   phase 2 (per switching lane, whole warp): the 32 lanes evaluate att_i for i = lane, lane+32, ..
//           in parallel, a shuffle scan builds the CDF, a ballot finds the first i with cs >= u2;
//           any |cs-u2| within the band up to that i -> FP64.
//   FP64 (rare, whole warp): lanes evaluate att_i in Float64 in parallel into shared memory, lane 0
//           performs the reference's sequential sums (abm.jl:123-130) so the result is the oracle's.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// att scratch: one double[Lpad] per warp in shared memory
__device__ __forceinline__ int warp_switch_decision(bool valid, float x, float y, int st, int follow, float u1, float u2,
                                                    const float* infx, const float* infy, const float* frac,
                                                    const double* dinfx, const double* dinfy, const double* dfrac,
                                                    int L, float eta, float dt, double deta, double ddt, float band,
                                                    double* att_scratch, unsigned int* nfallback) {
  const int lane = threadIdx.x & 31;
  // ---- phase 1
  float lam = 0.f;
  for (int i = 0; i < L; i++) {
    float d0 = x - infx[i], d1 = y - infy[i];
    float g2 = (float)st * frac[i];
    if (g2 < 0.1f) g2 = 0.1f;
    lam += eta * w_of_r2(fmaf(d0, d0, d1 * d1)) * g2;
  }
  int status = 0;  // 0 stay, 1 switch (FP32-decided), 2 needs FP64
  if (valid && lam == lam) {  // NaN lock (Q5): stay
    float thr = -expm1f(-lam * dt);
    float gap = u1 - thr;
    if (fabsf(gap) <= band * fmaxf(thr, 1e-3f)) status = 2;
    else if (gap < 0.f) status = 1;
  }
  int result = follow;
  // ---- phase 2: cooperative CDF walk for the FP32-decided switchers
  unsigned sw = __ballot_sync(0xffffffffu, status == 1);
  while (sw) {
    const int src = __ffs(sw) - 1;
    sw &= sw - 1;
    const float sx = __shfl_sync(0xffffffffu, x, src), sy = __shfl_sync(0xffffffffu, y, src);
    const float sst = (float)__shfl_sync(0xffffffffu, st, src);
    const float sinv = 1.0f / __shfl_sync(0xffffffffu, lam, src), su2 = __shfl_sync(0xffffffffu, u2, src);
    float carry = 0.f;
    int found = -1;
    bool fuzzy = false;
    for (int i0 = 0; i0 < L && found < 0 && !fuzzy; i0 += 32) {
      const int i = i0 + lane;
      float a = 0.f;
      if (i < L) {
        float d0 = sx - infx[i], d1 = sy - infy[i];
        float g2 = sst * frac[i];
        if (g2 < 0.1f) g2 = 0.1f;
        a = eta * w_of_r2(fmaf(d0, d0, d1 * d1)) * g2 * sinv;
      }
      const float cs = carry + warp_incl_scan(a, lane);
      const float g = cs - su2;
      const bool last = i == L - 1;                      // the clamp index (Q7) needs no comparison
      const unsigned hit = __ballot_sync(0xffffffffu, i < L && (g >= 0.f || last));
      const unsigned near = __ballot_sync(0xffffffffu, i < L && !last && fabsf(g) <= band);
      if (hit) {
        const int h = __ffs(hit) - 1;
        found = i0 + h;
        if (near & ((2u << h) - 1u)) fuzzy = true;       // a borderline comparison at or before the hit
      } else if (near) {
        fuzzy = true;
      }
      carry = __shfl_sync(0xffffffffu, cs, 31);
    }
    if (lane == src) {
      if (fuzzy || found < 0) status = 2;
      else result = found;
    }
  }
  // ---- FP64 re-evaluation (guard band), cooperative
  unsigned fb = __ballot_sync(0xffffffffu, status == 2);
  while (fb) {
    const int src = __ffs(fb) - 1;
    fb &= fb - 1;
    const double sx = (double)__shfl_sync(0xffffffffu, x, src), sy = (double)__shfl_sync(0xffffffffu, y, src);
    const double sst = (double)__shfl_sync(0xffffffffu, st, src);
    for (int i = lane; i < L; i += 32) {
      double d0 = sx - dinfx[i], d1 = sy - dinfy[i];
      double dist = exp(-sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1))));
      double g2 = __dmul_rn(sst, dfrac[i]);
      if (g2 < SMM_THETA) g2 = SMM_THETA;
      att_scratch[i] = __dmul_rn(__dmul_rn(deta, dist), g2);
    }
    __syncwarp();
    if (lane == src) {
      double lamd = 0.0;
      for (int i = 0; i < L; i++) lamd = __dadd_rn(lamd, att_scratch[i]);
      double thr = 1.0 - exp(-__dmul_rn(lamd, ddt));
      if ((double)u1 < thr) {
        double cs = 0.0;
        int k = L - 1;
        for (int i = 0; i < L; i++) {
          cs = __dadd_rn(cs, att_scratch[i] / lamd);
          if (!(cs < (double)u2)) { k = i; break; }
        }
        result = k;
      } else {
        result = follow;
      }
      (*nfallback)++;
    }
    __syncwarp();
  }
  return result;
}

// Variant of the warp decision whose cost does not depend on how many agents switch: phase 1 keeps
// every second running sum lam_i in shared memory (pref[(L+1)/2][blockDim.x], conflict-free), phase 2
// is a binary search over them plus one re-evaluated weight.  tab: float4 per influencer
// {x, y, eta*max(frac,0.1), eta*max(-frac,0.1)} (a NaN fraction stays NaN: Q5).
__device__ __forceinline__ int warp_switch_decision_prefix(bool valid, float x, float y, int st, int follow, float u1,
                                                           float u2, const float4* tab, const double* dinfx,
                                                           const double* dinfy, const double* dfrac, int L, float dt,
                                                           double deta, double ddt, float band, float* pref,
                                                           double* att_scratch, unsigned int* nfallback) {
  const int lane = threadIdx.x & 31, nt = blockDim.x;
  float* mypref = pref + threadIdx.x;
  const int M2 = (L + 1) >> 1;
  float lam = 0.f;
#pragma unroll 4
  for (int i = 0; i < L; i++) {  // abm.jl:103-123
    const float4 t = tab[i];
    const float d0 = x - t.x, d1 = y - t.y;
    lam = fmaf(w_of_r2(fmaf(d0, d0, d1 * d1)), st == 1 ? t.z : t.w, lam);
    if ((i & 1) || i == L - 1) mypref[(i >> 1) * nt] = lam;
  }
  int status = 0;  // 0 stay, 1 switch (FP32-decided), 2 needs FP64
  if (valid && lam == lam) {  // NaN lock (Q5): stay
    const float thr = -expm1f(-lam * dt);
    const float gap = u1 - thr;
    if (fabsf(gap) <= band * fmaxf(thr, 1e-3f)) status = 2;
    else if (gap < 0.f) status = 1;
  }
  int result = follow;
  {  // smallest k with lam_k / lam >= u2 (abm.jl:125-130)
    const float target = u2 * lam;
    int lo = 0, hi = M2 - 1;  // the last entry is the clamp (Q7)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (mypref[mid * nt] >= target) hi = mid; else lo = mid + 1;
    }
    const int i0 = 2 * lo;
    const float base = lo ? mypref[(lo - 1) * nt] : 0.f, top = mypref[lo * nt];
    const float4 t = tab[i0];
    const float d0 = x - t.x, d1 = y - t.y;
    const float cs0 = fmaf(w_of_r2(fmaf(d0, d0, d1 * d1)), st == 1 ? t.z : t.w, base);  // == phase-1 value
    const int k = (i0 + 1 < L && !(cs0 >= target)) ? i0 + 1 : i0;
    if (status == 1) {
      const float tol = band * lam;  // borderline comparisons at the decision boundary -> FP64
      const float pk = k == i0 ? cs0 : top, pkm1 = k == i0 ? base : cs0;
      const bool fz = (k < L - 1 && fabsf(pk - target) <= tol) || (k > 0 && fabsf(pkm1 - target) <= tol);
      if (fz) status = 2; else result = k;
    }
  }
  unsigned fb = __ballot_sync(0xffffffffu, status == 2);
  while (fb) {  // FP64 re-evaluation (guard band), cooperative
    const int src = __ffs(fb) - 1;
    fb &= fb - 1;
    const double sx = (double)__shfl_sync(0xffffffffu, x, src), sy = (double)__shfl_sync(0xffffffffu, y, src);
    const double sst = (double)__shfl_sync(0xffffffffu, st, src);
    for (int i = lane; i < L; i += 32) {
      double d0 = sx - dinfx[i], d1 = sy - dinfy[i];
      double dist = exp(-sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1))));
      double g2 = __dmul_rn(sst, dfrac[i]);
      if (g2 < SMM_THETA) g2 = SMM_THETA;
      att_scratch[i] = __dmul_rn(__dmul_rn(deta, dist), g2);
    }
    __syncwarp();
    if (lane == src) {
      double lamd = 0.0;
      for (int i = 0; i < L; i++) lamd = __dadd_rn(lamd, att_scratch[i]);
      double thr = 1.0 - exp(-__dmul_rn(lamd, ddt));
      if ((double)u1 < thr) {
        double cs = 0.0;
        int k = L - 1;
        for (int i = 0; i < L; i++) {
          cs = __dadd_rn(cs, att_scratch[i] / lamd);
          if (!(cs < (double)u2)) { k = i; break; }
        }
        result = k;
      } else {
        result = follow;
      }
      (*nfallback)++;
    }
    __syncwarp();
  }
  return result;
}
