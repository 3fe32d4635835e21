// fft_generic.cuh -- shared-memory mixed-radix (3, 2^2, 2) FFT passes used by fft_conv.cu (one kernel per pass group)
// and pde_cluster.cu (the whole convolution inside one thread-block cluster).  Any block size; the functions contain
// __syncthreads() and must be called by every thread of the CTA.
#pragma once
#include "smm_common.cuh"

namespace {

template <typename real2> __device__ __forceinline__ real2 cadd(real2 a, real2 b) { a.x += b.x; a.y += b.y; return a; }
template <typename real2> __device__ __forceinline__ real2 csub(real2 a, real2 b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename real2> __device__ __forceinline__ real2 cmul(real2 a, real2 b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename real2> __device__ __forceinline__ real2 cmulc(real2 a, real2 b) { return mk2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)

// FFT length P = r3 * m with m = 2^a and r3 in {1, 3}: P = 3*2^a avoids padding 2N-1 up to the next
// power of two (705 -> 768 instead of 1024).  A radix-3 pass is the first pass of the forward
// (decimation-in-frequency) transform and the last of the inverse; between them each third is an
// independent power-of-two FFT.  Spectrum order after the forward transform ("permuted order"):
//   position(k) = (k mod r3) * m + bitrev_a(k div r3)
struct FftDims { int P, m, a, r3; };

// Shared-memory swizzle of an FFT column: logical element i lives at physical slot i ^ ((i >> 2) & 15).
// The radix-2^2 passes with small butterfly distance (q = 1, 4: element strides 4 and 16) were 4-way
// bank-conflicted in the natural layout (48 % of all shared wavefronts of fft_cols at P = 768); with the
// swizzle every pass of every supported length is conflict-free (checked by enumeration).
// It permutes within aligned blocks of 64 (of 8 inside a trailing half block), so it is a bijection on
// [0, P) for every P = r3 * 2^a, a >= 5.
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 2) & 15); }

__host__ __device__ __forceinline__ int brev_a(int i, int a) {
#if defined(__CUDA_ARCH__)
  return a ? (int)(__brev((unsigned)i) >> (32 - a)) : 0;
#else
  int r = 0;
  for (int b = 0; b < a; b++) r |= ((i >> b) & 1) << (a - 1 - b);
  return r;
#endif
}
__host__ __device__ __forceinline__ int pos_of_freq(int k, FftDims d) {
  return d.r3 == 3 ? (k % 3) * d.m + brev_a(k / 3, d.a) : brev_a(k, d.a);
}
__device__ __forceinline__ int freq_of_pos(int p, FftDims d) {
  if (d.r3 == 3) { const int r = p / d.m; return 3 * brev_a(p - r * d.m, d.a) + r; }
  return brev_a(p, d.a);
}

// Twiddle table (P entries): [0, m) stage-major power-of-two twiddles tw[half + j] = exp(-2 pi i j/(2 half))
// (consecutive lanes read consecutive words; the strided form tw[j * m/2^s] was up to 32-way
// bank-conflicted), and for r3 = 3 the radix-3 twiddles tw[m + j] = exp(-2 pi i j / P), j < 2m.
template <typename real>
__global__ void twiddle_kernel(FftDims d, typename vec2<real>::type* tw) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.P) return;
  double s, c;
  if (e < d.m) {
    if (e == 0) { tw[0] = mk2((real)1, (real)0); return; }
    int half = 1 << (31 - __clz(e));
    sincospi(-(double)(e - half) / (double)half, &s, &c);  // exp(-2 pi i j / (2 half))
  } else {
    sincospi(-2.0 * (double)(e - d.m) / (double)d.P, &s, &c);
  }
  tw[e] = mk2((real)c, (real)s);
}

#define SMM_W3_RE (-0.5)
#define SMM_W3_IM (-0.86602540378443864676)  // w3 = exp(-2 pi i / 3)

// ncol independent length-P FFTs in shared memory (column stride cs), decimation in frequency:
// natural-order input, permuted-order output.  Power-of-two part in radix-2^2 passes (two butterfly
// stages per pass in registers: half the shared-memory traffic and barriers of radix 2).
template <typename real>
__device__ __forceinline__ void fft_dif_fwd(typename vec2<real>::type* sh, int ncol, int cs, FftDims d,
                            const typename vec2<real>::type* tw) {
  using real2 = typename vec2<real>::type;
  const int tid = threadIdx.x, nt = blockDim.x, m = d.m, a = d.a, r3 = d.r3;
  if (r3 == 3) {
    const real2* t3 = tw + m;
    for (int t = tid; t < ncol * m; t += nt) {
      const int c = t >> a, j = t & (m - 1);
      real2* col = sh + c * cs;
      const int i0 = swz(j), i1 = swz(j + m), i2 = swz(j + 2 * m);
      const real2 a0 = col[i0], a1 = col[i1], a2 = col[i2];
      // y1 = a0 + w3 a1 + w3^2 a2, y2 = a0 + w3^2 a1 + w3 a2  (w3^2 = conj(w3), w3^4 = w3)
      const real sr = a1.x + a2.x, si = a1.y + a2.y, dr = a1.x - a2.x, di = a1.y - a2.y;
      const real tr = a0.x + (real)SMM_W3_RE * sr, ti = a0.y + (real)SMM_W3_RE * si;
      const real ur = -(real)SMM_W3_IM * di, ui = (real)SMM_W3_IM * dr;  // i*Im(w3)*(a1-a2) = (-Im*di, Im*dr)
      col[i0] = mk2(a0.x + sr, a0.y + si);
      col[i1] = cmul(mk2(tr + ur, ti + ui), t3[j]);
      col[i2] = cmul(mk2(tr - ur, ti - ui), t3[2 * j]);
    }
    __syncthreads();
  }
  const int nsub = ncol * r3, qp = m >> 2;
  int s = a;
  for (; s >= 2; s -= 2) {
    const int q = 1 << (s - 2);
    for (int t = tid; t < nsub * qp; t += nt) {
      const int cc = t >> (a - 2), bt = t & (qp - 1);
      const int c = r3 == 3 ? cc / 3 : cc, r = cc - c * r3;
      const int j = bt & (q - 1), i0 = ((bt >> (s - 2)) << s) + j;
      real2* col = sh + c * cs;
      const int e0 = r * m + i0, p0 = swz(e0), p1 = swz(e0 + q), p2 = swz(e0 + 2 * q), p3 = swz(e0 + 3 * q);
      const real2 a0 = col[p0], a1 = col[p1], a2 = col[p2], a3 = col[p3];
      const real2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), tw[2 * q + j]);
      const real2 b1 = cadd(a1, a3), b3 = cmul(csub(a1, a3), tw[3 * q + j]);
      const real2 w = tw[q + j];
      col[p0] = cadd(b0, b1);
      col[p1] = cmul(csub(b0, b1), w);
      col[p2] = cadd(b2, b3);
      col[p3] = cmul(csub(b2, b3), w);
    }
    __syncthreads();
  }
  if (s == 1) {  // odd a: last radix-2 stage (twiddle 1)
    for (int t = tid; t < nsub * (m >> 1); t += nt) {
      const int cc = t >> (a - 1), bt = t & ((m >> 1) - 1);
      const int c = r3 == 3 ? cc / 3 : cc, r = cc - c * r3;
      real2* col = sh + c * cs;
      const int p0 = swz(r * m + 2 * bt), p1 = swz(r * m + 2 * bt + 1);
      const real2 p = col[p0], q2 = col[p1];
      col[p0] = cadd(p, q2);
      col[p1] = csub(p, q2);
    }
    __syncthreads();
  }
}

// decimation in time with conjugate twiddles: permuted-order input, natural-order output (unscaled inverse)
template <typename real>
__device__ __forceinline__ void fft_dit_inv(typename vec2<real>::type* sh, int ncol, int cs, FftDims d,
                            const typename vec2<real>::type* tw) {
  using real2 = typename vec2<real>::type;
  const int tid = threadIdx.x, nt = blockDim.x, m = d.m, a = d.a, r3 = d.r3;
  const int nsub = ncol * r3, qp = m >> 2;
  int s = 1;
  if (a & 1) {  // odd a: first radix-2 stage (twiddle 1)
    for (int t = tid; t < nsub * (m >> 1); t += nt) {
      const int cc = t >> (a - 1), bt = t & ((m >> 1) - 1);
      const int c = r3 == 3 ? cc / 3 : cc, r = cc - c * r3;
      real2* col = sh + c * cs;
      const int p0 = swz(r * m + 2 * bt), p1 = swz(r * m + 2 * bt + 1);
      const real2 p = col[p0], q2 = col[p1];
      col[p0] = cadd(p, q2);
      col[p1] = csub(p, q2);
    }
    __syncthreads();
    s = 2;
  }
  for (; s + 1 <= a; s += 2) {  // stages s and s+1
    const int q = 1 << (s - 1);
    for (int t = tid; t < nsub * qp; t += nt) {
      const int cc = t >> (a - 2), bt = t & (qp - 1);
      const int c = r3 == 3 ? cc / 3 : cc, r = cc - c * r3;
      const int j = bt & (q - 1), i0 = ((bt >> (s - 1)) << (s + 1)) + j;
      real2* col = sh + c * cs;
      const int e0 = r * m + i0, p0 = swz(e0), p1 = swz(e0 + q), p2 = swz(e0 + 2 * q), p3 = swz(e0 + 3 * q);
      const real2 a0 = col[p0], a2 = col[p2];
      const real2 w1 = tw[q + j];
      const real2 t1 = cmulc(col[p1], w1), t3 = cmulc(col[p3], w1);
      const real2 b0 = cadd(a0, t1), b1 = csub(a0, t1), b2 = cadd(a2, t3), b3 = csub(a2, t3);
      const real2 u2 = cmulc(b2, tw[2 * q + j]), u3 = cmulc(b3, tw[3 * q + j]);
      col[p0] = cadd(b0, u2);
      col[p2] = csub(b0, u2);
      col[p1] = cadd(b1, u3);
      col[p3] = csub(b1, u3);
    }
    __syncthreads();
  }
  if (r3 == 3) {
    const real2* t3 = tw + m;
    for (int t = tid; t < ncol * m; t += nt) {
      const int c = t >> a, j = t & (m - 1);
      real2* col = sh + c * cs;
      const int i0 = swz(j), i1 = swz(j + m), i2 = swz(j + 2 * m);
      const real2 b0 = col[i0], b1 = cmulc(col[i1], t3[j]), b2 = cmulc(col[i2], t3[2 * j]);
      // a1 = b0 + conj(w3) b1 + w3 b2, a2 = b0 + w3 b1 + conj(w3) b2
      const real sr = b1.x + b2.x, si = b1.y + b2.y, dr = b1.x - b2.x, di = b1.y - b2.y;
      const real tr = b0.x + (real)SMM_W3_RE * sr, ti = b0.y + (real)SMM_W3_RE * si;
      const real ur = -(real)SMM_W3_IM * di, ui = (real)SMM_W3_IM * dr;
      col[i0] = mk2(b0.x + sr, b0.y + si);
      col[i1] = mk2(tr - ur, ti - ui);
      col[i2] = mk2(tr + ur, ti + ui);
    }
    __syncthreads();
  }
}

}  // namespace
