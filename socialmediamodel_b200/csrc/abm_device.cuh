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
