// fft_conv.cu -- K6 (scalable form): the interaction force F = (w*(x'-x)rho)/(w*rho) of
// pde.jl:49-59 as a zero-padded FFT correlation.  The reference multiplies by the dense
// K_matrix/W_matrix (O(N^2) per right-hand side, 248 GB at 353x353); the operator is block
// Toeplitz, so three P x P circular convolutions (P = 2^a or 3*2^a >= 2N-1) give the same sums in
// O(N log N).  Hand-written mixed-radix (3, 2^2, 2) shared-memory FFTs (no cuFFT):
//
//   fft_rows_fwd   two real rows per complex FFT (DIF, natural -> bit-reversed), split into the two
//                  half spectra                                       rho[b][iy][ix] -> A[b][iy][kx]
//   fft_cols       per group of COLS kx-columns: DIF forward over y, multiply by the three kernel
//                  spectra (stored bit-reversed in ky, pre-scaled by 1/P^2), DIT inverse, keep
//                  iy < Ny                                              A -> G[k][b][iy][kx], k=0..2
//   fft_rows_inv   Hermitian-extend, two real rows per complex DIT inverse, keep ix < Nx, divide
//                                                                        G -> Fx, Fy [b][iy][ix]
// The kernel spectra are computed once per grid with the same kernels instantiated in Float64.
#include <stdlib.h>

#include "smm_internal.h"

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
__device__ void fft_dif_fwd(typename vec2<real>::type* sh, int ncol, int cs, FftDims d,
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
__device__ void fft_dit_inv(typename vec2<real>::type* sh, int ncol, int cs, FftDims d,
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

// circular kernel images c_k[dy mod P][dx mod P] (Float64): w, -w*hx*dx, -w*hy*dy  (pde.jl:14-20 with
// the correlation written as a convolution: out[i] = sum_j T[j-i] rho[j] = sum_j T~[i-j] rho[j])
__global__ void kernel_image_kernel(int Nx, int Ny, int P, double hx, double hy, double* img) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)P * P) return;
  int py = (int)(e / P), px = (int)(e % P);
  int dy = py < P / 2 ? py : py - P, dx = px < P / 2 ? px : px - P;
  double w = 0.0;
  if (abs(dx) <= Nx - 1 && abs(dy) <= Ny - 1) {
    double ddx = hx * dx, ddy = hy * dy;
    w = exp(-sqrt(ddx * ddx + ddy * ddy));
  }
  img[e] = w;
  img[(size_t)P * P + e] = -w * hx * dx;
  img[2 * (size_t)P * P + e] = -w * hy * dy;
}

// rows forward: in[batch][nrows][rowlen] real -> out[batch][nrows][P/2+1] complex.  CTA = one row pair.
template <typename real>
__global__ void __launch_bounds__(256) fft_rows_fwd_kernel(int nrows, int rowlen, int rowpitch, FftDims d, const real* __restrict__ in,
                                    typename vec2<real>::type* __restrict__ out,
                                    const typename vec2<real>::type* __restrict__ twg) {
  using real2 = typename vec2<real>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int P = d.P;
  real2* sh = reinterpret_cast<real2*>(smem);
  real2* tw = sh + P;
  const int b = blockIdx.y, ra = blockIdx.x * 2, rb = ra + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  const real* rowa = in + ((size_t)b * nrows + ra) * rowpitch;
  const real* rowb = in + ((size_t)b * nrows + rb) * rowpitch;
  const bool hasb = rb < nrows;
  for (int i = tid; i < P; i += nt) tw[i] = twg[i];  // constant table: before the dependency wait
  pdl_wait();
  pdl_launch_dependents();
  for (int i = tid; i < P; i += nt) sh[swz(i)] = i < rowlen ? mk2(rowa[i], hasb ? rowb[i] : (real)0) : mk2((real)0, (real)0);
  __syncthreads();
  fft_dif_fwd<real>(sh, 1, P, d, tw);
  const int hp1 = P / 2 + 1;
  real2* oa = out + ((size_t)b * nrows + ra) * hp1;
  real2* ob = out + ((size_t)b * nrows + rb) * hp1;
  for (int k = tid; k < hp1; k += nt) {
    const real2 zk = sh[swz(pos_of_freq(k, d))], zm = sh[swz(pos_of_freq(k ? P - k : 0, d))];
    oa[k] = mk2((real)0.5 * (zk.x + zm.x), (real)0.5 * (zk.y - zm.y));
    if (hasb) ob[k] = mk2((real)0.5 * (zk.y + zm.y), (real)-0.5 * (zk.x - zm.x));
  }
}

// columns.  MODE 0: convolution (forward, x3 kernel spectra, inverse, keep iy < nkeep).
//           MODE 1: spectrum only: out[k=blockIdx.y][kx][ky bit-reversed] scaled by `scale`.
template <typename real, int MODE>
__global__ void __launch_bounds__(512) fft_cols_kernel(int nrows, int nkeep, FftDims d, int cols, double scale,
                                const typename vec2<real>::type* __restrict__ A,
                                const typename vec2<real>::type* __restrict__ khat,
                                typename vec2<real>::type* __restrict__ G,
                                const typename vec2<real>::type* __restrict__ twg, size_t gstride) {
  using real2 = typename vec2<real>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int P = d.P;
  const int cs = P + 1;
  real2* R = reinterpret_cast<real2*>(smem);       // [cols][cs]
  real2* Wk = R + (size_t)cols * cs;               // [cols][cs] (MODE 0)
  real2* tw = MODE == 0 ? Wk + (size_t)cols * cs : Wk;
  const int hp1 = P / 2 + 1;
  const int b = blockIdx.y, kx0 = blockIdx.x * cols;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < P; i += nt) tw[i] = twg[i];  // constant table: before the dependency wait
  pdl_wait();
  pdl_launch_dependents();
  const real2* Ab = A + (size_t)b * nrows * hp1;
  for (int e = tid; e < P * cols; e += nt) {
    const int jy = e / cols, c = e - jy * cols;
    real2 v = mk2((real)0, (real)0);
    if (jy < nrows && kx0 + c < hp1) v = Ab[(size_t)jy * hp1 + kx0 + c];
    R[c * cs + swz(jy)] = v;
  }
  __syncthreads();
  fft_dif_fwd<real>(R, cols, cs, d, tw);
  if (MODE == 1) {
    for (int e = tid; e < P * cols; e += nt) {
      const int c = e / P, i = e - c * P;
      if (kx0 + c < hp1) {
        real2 v = R[c * cs + swz(i)];
        G[((size_t)b * hp1 + kx0 + c) * P + i] = mk2((real)(v.x * scale), (real)(v.y * scale));
      }
    }
    return;
  }
  for (int k = 0; k < 3; k++) {
    for (int e = tid; e < P * cols; e += nt) {
      const int c = e / P, i = e - c * P;
      real2 v = mk2((real)0, (real)0);
      if (kx0 + c < hp1) v = cmul(R[c * cs + swz(i)], khat[((size_t)k * hp1 + kx0 + c) * P + i]);
      Wk[c * cs + swz(i)] = v;
    }
    __syncthreads();
    fft_dit_inv<real>(Wk, cols, cs, d, tw);
    real2* Gk = G + (size_t)k * gstride + (size_t)b * nkeep * hp1;
    for (int e = tid; e < nkeep * cols; e += nt) {
      const int iy = e / cols, c = e - iy * cols;
      if (kx0 + c < hp1) Gk[(size_t)iy * hp1 + kx0 + c] = Wk[c * cs + swz(iy)];
    }
    __syncthreads();
  }
}

// rows inverse + division: G[k][b][iy][kx] -> Fx, Fy [b][iy][ix].  CTA = one row pair, 3 FFTs.
template <typename real>
__global__ void __launch_bounds__(256) fft_rows_inv_kernel(int Nx, int Ny, int Pw, FftDims d, const typename vec2<real>::type* __restrict__ G,
                                    size_t gstride, real* __restrict__ Fx, real* __restrict__ Fy,
                                    const typename vec2<real>::type* __restrict__ twg) {
  using real2 = typename vec2<real>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int P = d.P;
  const int cs = P + 1;
  real2* sh = reinterpret_cast<real2*>(smem);  // [3][cs]
  real2* tw = sh + 3 * cs;
  const int hp1 = P / 2 + 1;
  const int b = blockIdx.y, ra = blockIdx.x * 2, rb = ra + 1;
  const bool hasb = rb < Ny;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < P; i += nt) tw[i] = twg[i];  // constant table: before the dependency wait
  pdl_wait();
  pdl_launch_dependents();
  // the DIT inverse wants its input bit-reversed: iterate over shared-memory positions (conflict-free
  // stores) and fetch spectrum index brev(pos) from the 4 KB global row (L1-resident) -- scattering
  // into bit-reversed shared addresses instead was a 32-way bank conflict
  for (int e = tid; e < 3 * P; e += nt) {
    const int k = e / P, pos = e - k * P;
    const int kk = freq_of_pos(pos, d);
    const bool up = kk > P / 2;
    const int src = up ? P - kk : kk;
    const real2* Gk = G + (size_t)k * gstride + (size_t)b * Ny * hp1;
    real2 a = Gk[(size_t)ra * hp1 + src];
    real2 bb = hasb ? Gk[(size_t)rb * hp1 + src] : mk2((real)0, (real)0);
    if (up) { a.y = -a.y; bb.y = -bb.y; }
    sh[k * cs + swz(pos)] = mk2(a.x - bb.y, a.y + bb.x);  // A + i B
  }
  __syncthreads();
  fft_dit_inv<real>(sh, 3, cs, d, tw);
  for (int ix = tid; ix < Nx; ix += nt) {
    const int px = swz(ix);
    const real2 o0 = sh[px], ox = sh[cs + px], oy = sh[2 * cs + px];
    const size_t oa = ((size_t)b * Ny + ra) * Pw + ix;
    Fx[oa] = ox.x / o0.x;  // pde.jl:58 force ./ norms
    Fy[oa] = oy.x / o0.x;
    if (hasb) {
      Fx[oa + Pw] = ox.y / o0.y;
      Fy[oa + Pw] = oy.y / o0.y;
    }
  }
}

template <typename real> int cols_for(int P) {
  int c = (int)(32768 / ((size_t)P * 2 * sizeof(real)));  // <= 32 KB per buffer: 3 CTAs per SM at P = 1024
  return c < 1 ? 1 : (c > 8 ? 8 : c);
}

}  // namespace

struct smm_fft_plan {
  int Nx, Ny, Pw, P, B, precision;  // Pw: row pitch of rhosum / Fx / Fy
  smm_fft_fast* fast;               // register-resident FP32 path (fft_fast.cu); null = the kernels of this file
  FftDims dims;
  void* tw;    // real2 [P], stage-major
  void* khat;  // real2 [3][P/2+1][P]
  void* ahat;  // real2 [B][Ny][P/2+1]
  void* g;     // real2 [3][B][Ny][P/2+1]
};

void smm_fft_plan_destroy(smm_fft_plan* p) {
  if (!p) return;
  void* ptrs[] = {p->tw, p->khat, p->ahat, p->g};
  for (void* q : ptrs) if (q) cudaFree(q);
  smm_fft_fast_destroy(p->fast);
  delete p;
}

template <typename real>
static int set_smem(smm_ctx* ctx, int P, int cols) {
  using real2 = typename vec2<real>::type;
  size_t sm_cols = ((size_t)2 * cols * (P + 1) + P) * sizeof(real2);
  size_t sm_cols1 = ((size_t)cols * (P + 1) + P) * sizeof(real2);
  size_t sm_rows = ((size_t)P + P) * sizeof(real2);
  size_t sm_inv = ((size_t)3 * (P + 1) + P) * sizeof(real2);
  SMM_CUDA(ctx, cudaFuncSetAttribute(fft_cols_kernel<real, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cols));
  SMM_CUDA(ctx, cudaFuncSetAttribute(fft_cols_kernel<real, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cols1));
  SMM_CUDA(ctx, cudaFuncSetAttribute(fft_rows_fwd_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_rows));
  SMM_CUDA(ctx, cudaFuncSetAttribute(fft_rows_inv_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_inv));
  return SMM_OK;
}

int smm_fft_plan_create(smm_ctx* ctx, int Nx, int Ny, int Pw, double hx, double hy, int B, int precision, smm_fft_plan** out) {
  *out = nullptr;
  // smallest P = 2^a or 3*2^a (a >= 5) with P >= 2N-1
  const int need = 2 * (Nx > Ny ? Nx : Ny) - 1;
  FftDims d = {0, 0, 0, 1};
  for (int a = 5; a <= 11 && !d.P; a++) {
    const int p2 = 1 << a, p3 = 3 << (a - 1);  // ... 2^a  and  3*2^(a-1) < 2^(a+1)
    if (p2 >= need) d = {p2, p2, a, 1};
    else if (a - 1 >= 5 && p3 >= need) d = {p3, 1 << (a - 1), a - 1, 3};
  }
  {
    const char* ev = getenv("SMM_FFT_POW2");  // debugging: force power-of-two lengths
    if (ev && ev[0] == '1' && d.r3 == 3) d = {4 * d.m, 4 * d.m, d.a + 2, 1};
  }
  const int P = d.P;
  SMM_REQUIRE(ctx, P > 0 && P <= 2048, "FFT convolution: grid %dx%d needs P=%d > 2048", Nx, Ny, P);
  smm_fft_plan* p = new smm_fft_plan();
  p->Nx = Nx; p->Ny = Ny; p->Pw = Pw; p->P = P; p->dims = d; p->B = B; p->precision = precision;
  p->tw = p->khat = p->ahat = p->g = nullptr;
  p->fast = nullptr;
  const size_t c2 = precision == SMM_PREC_F64 ? 16 : 8;
  const int hp1 = P / 2 + 1;
  // FP32 and P = 96 / 192 / 384 / 768 (the grids 47, 89, 177, 353): the register-resident path of fft_fast.cu for
  // batches with enough rows to fill the GPU (measured on B200: 16 x 177^2 0.415 -> 0.336 ms, 64 x 353^2 5.48 -> 3.80 ms per
  // RK4 step; a single 89^2 instance, 13 column CTAs, is latency-bound and 40 % slower than the kernels of this file,
  // which spread one instance over 129 CTAs).  SMM_FFT_FAST=1 / 0 forces the choice; the kernels of this file also serve
  // Float64 and every other length.
  bool want_fast = precision != SMM_PREC_F64 && smm_fft_fast_supported(P) && (size_t)B * Ny >= 1024;
  {
    const char* ev = getenv("SMM_FFT_FAST");
    if (ev && ev[0] == '0') want_fast = false;
    if (ev && ev[0] == '1') want_fast = precision != SMM_PREC_F64 && smm_fft_fast_supported(P);
  }
#define PL_CUDA(expr)                                                                                    \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      smm_fft_plan_destroy(p);                                                                           \
      for (void* t : tmp) if (t) cudaFree(t);                                                            \
      cudaGetLastError();                                                                                \
      return smm_fail(ctx, e__ == cudaErrorMemoryAllocation ? SMM_ERR_NOMEM : SMM_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(e__)); \
    }                                                                                                    \
  } while (0)
  void* tmp[4] = {nullptr, nullptr, nullptr, nullptr};  // img, twd, ahat_d, khat_d
  PL_CUDA(cudaMalloc(&p->tw, (size_t)P * c2));
  PL_CUDA(cudaMalloc(&p->khat, (size_t)3 * hp1 * P * c2));
  if (!want_fast) {
    PL_CUDA(cudaMalloc(&p->ahat, (size_t)B * Ny * hp1 * c2));
    PL_CUDA(cudaMalloc(&p->g, (size_t)3 * B * Ny * hp1 * c2));
  }
  PL_CUDA(cudaMalloc(&tmp[0], (size_t)3 * P * P * 8));
  PL_CUDA(cudaMalloc(&tmp[1], (size_t)P * 16));
  PL_CUDA(cudaMalloc(&tmp[2], (size_t)3 * P * hp1 * 16));
  if (precision != SMM_PREC_F64) PL_CUDA(cudaMalloc(&tmp[3], (size_t)3 * hp1 * P * 16));
  {
    int rc = set_smem<double>(ctx, P, cols_for<double>(P));
    if (rc == SMM_OK && precision != SMM_PREC_F64) rc = set_smem<float>(ctx, P, cols_for<float>(P));
    if (rc != SMM_OK) { smm_fft_plan_destroy(p); for (void* t : tmp) if (t) cudaFree(t); return rc; }
  }
  // twiddles (Float64 and working precision), kernel images, their spectra in Float64
  twiddle_kernel<double><<<(P + 255) / 256, 256, 0, ctx->stream>>>(d, (double2*)tmp[1]);
  if (precision == SMM_PREC_F64) twiddle_kernel<double><<<(P + 255) / 256, 256, 0, ctx->stream>>>(d, (double2*)p->tw);
  else twiddle_kernel<float><<<(P + 255) / 256, 256, 0, ctx->stream>>>(d, (float2*)p->tw);
  kernel_image_kernel<<<(unsigned)(((size_t)P * P + 255) / 256), 256, 0, ctx->stream>>>(Nx, Ny, P, hx, hy, (double*)tmp[0]);
  {
    const int cols = cols_for<double>(P);
    size_t sm_rows = ((size_t)P + P) * 16;
    fft_rows_fwd_kernel<double><<<dim3(P / 2, 3), 256, sm_rows, ctx->stream>>>(P, P, P, d, (const double*)tmp[0], (double2*)tmp[2], (const double2*)tmp[1]);
    size_t sm_cols1 = ((size_t)cols * (P + 1) + P) * 16;
    double2* kd = precision == SMM_PREC_F64 ? (double2*)p->khat : (double2*)tmp[3];
    fft_cols_kernel<double, 1><<<dim3((hp1 + cols - 1) / cols, 3), 512, sm_cols1, ctx->stream>>>(
        P, 0, d, cols, 1.0 / ((double)P * (double)P), (const double2*)tmp[2], nullptr, kd, (const double2*)tmp[1], 0);
    ctx->launches += 5;
    PL_CUDA(cudaGetLastError());
    if (precision != SMM_PREC_F64) {
      int rc = smm_launch_convert_f64_to_real(ctx, (const double*)tmp[3], p->khat, (size_t)3 * hp1 * P * 2, precision);
      if (rc != SMM_OK) { smm_fft_plan_destroy(p); for (void* t : tmp) if (t) cudaFree(t); return rc; }
    }
    if (want_fast) {  // the same Float64 spectra, re-ordered for the fast path
      std::vector<int> slot(P);
      for (int k = 0; k < P; k++) slot[k] = pos_of_freq(k, d);
      int* dslot = nullptr;
      PL_CUDA(cudaMalloc((void**)&dslot, (size_t)P * sizeof(int)));
      cudaError_t e = cudaMemcpyAsync(dslot, slot.data(), (size_t)P * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
      int rc = e == cudaSuccess ? smm_fft_fast_create(ctx, P, Nx, Ny, Pw, B, tmp[3], dslot, &p->fast) : SMM_ERR_CUDA;
      cudaStreamSynchronize(ctx->stream);
      cudaFree(dslot);
      // the fast path keeps one real number per spectrum point (even kernel image: real spectrum; odd ones: imaginary);
      // the dropped parts must be rounding noise of the Float64 transform
      if (rc == SMM_OK && !(smm_fft_fast_residual(p->fast) < 1e-10))
        rc = smm_fail(ctx, SMM_ERR_UNSUPPORTED, "FFT convolution: kernel spectra are not real / imaginary (residual %g)", smm_fft_fast_residual(p->fast));
      if (rc != SMM_OK) { smm_fft_plan_destroy(p); for (void* t : tmp) if (t) cudaFree(t); return rc; }
    }
  }
  PL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (void* t : tmp) if (t) cudaFree(t);
#undef PL_CUDA
  *out = p;
  return SMM_OK;
}

template <typename real>
static int fft_conv_launch(smm_ctx* ctx, smm_fft_plan* p, const void* rhosum, void* Fx, void* Fy) {
  using real2 = typename vec2<real>::type;
  const int P = p->P, hp1 = P / 2 + 1;
  const FftDims d = p->dims;
  // columns per CTA: as many as shared memory allows for big batches (coalescing), fewer for small
  // problems so that the column kernel still fills the GPU (C3: one 89x89 instance -> 129 CTAs)
  int cols = cols_for<real>(P);
  while (cols > 1 && (size_t)((hp1 + cols - 1) / cols) * p->B < (size_t)2 * ctx->sm_count) cols >>= 1;
  const size_t gstride = (size_t)p->B * p->Ny * hp1;
  size_t sm_rows = ((size_t)P + P) * sizeof(real2);
  size_t sm_cols = ((size_t)2 * cols * (P + 1) + P) * sizeof(real2);
  size_t sm_inv = ((size_t)3 * (P + 1) + P) * sizeof(real2);
  const int rt = P >= 512 ? 256 : 128;
  static const bool pdl = !(getenv("SMM_PDE_PDL") && getenv("SMM_PDE_PDL")[0] == '0');
  SMM_CUDA(ctx, smm_launch_pdl(pdl ? ctx->sm_count : 0, fft_rows_fwd_kernel<real>, dim3((p->Ny + 1) / 2, p->B), dim3(rt), sm_rows, ctx->stream,
                               p->Ny, p->Nx, p->Pw, d, (const real*)rhosum, (real2*)p->ahat, (const real2*)p->tw));
  const int ct = (P * cols) / 2 >= 2048 ? 512 : ((P * cols) / 2 >= 256 ? 256 : 128);
  SMM_CUDA(ctx, smm_launch_pdl(pdl ? ctx->sm_count : 0, fft_cols_kernel<real, 0>, dim3((hp1 + cols - 1) / cols, p->B), dim3(ct), sm_cols, ctx->stream,
                               p->Ny, p->Ny, d, cols, 1.0, (const real2*)p->ahat, (const real2*)p->khat, (real2*)p->g,
                               (const real2*)p->tw, gstride));
  SMM_CUDA(ctx, smm_launch_pdl(pdl ? ctx->sm_count : 0, fft_rows_inv_kernel<real>, dim3((p->Ny + 1) / 2, p->B), dim3(rt), sm_inv, ctx->stream,
                               p->Nx, p->Ny, p->Pw, d, (const real2*)p->g, gstride, (real*)Fx, (real*)Fy, (const real2*)p->tw));
  ctx->launches += 3;
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

int smm_fft_conv(smm_ctx* ctx, smm_fft_plan* p, const void* rhosum, void* Fx, void* Fy) {
  if (p->fast) return smm_fft_fast_conv(ctx, p->fast, rhosum, Fx, Fy);
  return p->precision == SMM_PREC_F64 ? fft_conv_launch<double>(ctx, p, rhosum, Fx, Fy)
                                      : fft_conv_launch<float>(ctx, p, rhosum, Fx, Fy);
}
