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

#include "fft_generic.cuh"
#include "smm_internal.h"

namespace {

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

void smm_fft_plan_tables(const smm_fft_plan* p, int dims[4], const void** tw, const void** khat) {
  dims[0] = p->dims.P; dims[1] = p->dims.m; dims[2] = p->dims.a; dims[3] = p->dims.r3;
  *tw = p->tw; *khat = p->khat;
}

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
