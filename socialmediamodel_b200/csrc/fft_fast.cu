// fft_fast.cu -- K6, throughput form: the interaction force F = (w*(x'-x)rho)/(w*rho) of pde.jl:49-59 as a zero-padded
// FFT correlation with register-resident passes (fft12_core.cuh) for the FFT lengths P = 96, 192, 384, 768 that the
// grids 47, 89, 177 and 353 need, FP32.  Same three steps as fft_conv.cu (which stays as the Float64 / any-length path
// and as the cross-check), different machinery:
//
//   fast_rows_fwd  CTA = 4 PAIRS grid rows = 2 PAIRS complex FFTs (two real rows per FFT, two FFTs per thread in packed
//                  FP32x2 registers); pass-1 operands come straight from global memory; the two half spectra are
//                  separated from the digit-reversed result through one shared-memory exchange   rho -> A[row][kx]
//   fast_cols      CTA = 2 PAIRS adjacent spectrum columns of one instance; forward transform, the spectrum stays in
//                  registers, and for each of the three kernel spectra: scale (the spectra of the even / odd kernel
//                  images are purely real / purely imaginary: one real factor per point), transposed inverse whose
//                  first pass is thread-local, rows iy < Ny stored from registers   A -> G[k][row quad][kx][row mod 4]
//   fast_rows_inv  CTA = 4 PAIRS grid rows; Hermitian extension while loading pass-1 operands, inverse in
//                  decimation-in-frequency form (digit-reversed output read back through a position table), k = 0 first
//                  so that 1/(w*rho) waits in registers for the two force components            G -> Fx, Fy
//
// Work per 1-D transform: 12 points per thread in 3-4 passes (radix 12, then radix 4 / 2) instead of ten radix-2^2 half
// passes through shared memory; every exchange is conflict-free (tools/fft12_check.cpp enumerates them).
#include <stdlib.h>

#include <algorithm>

#include "fft12_core.cuh"
#include "smm_internal.h"

namespace {

using namespace f12;

template <int P, int PAIRS> struct Cfg {
  static constexpr int T = Plan<P>::T, NEX = Plan<P>::NEX, THREADS = T * PAIRS, ROWS = 4 * PAIRS, COLS = 2 * PAIRS;
  static constexpr size_t BUF = (size_t)P * PAIRS * sizeof(float4);
  static constexpr size_t SMEM = 2 * BUF + (size_t)P * sizeof(float2) + (size_t)P * sizeof(unsigned short);
  static constexpr int MINB = 512 / THREADS > 0 ? 512 / THREADS : 1;  // <= 512 resident threads per SM: 128 registers each
};

template <int P, int PAIRS> struct Smem {
  float4 *b0, *b1;
  float2* tw;
  unsigned short* postab;
  __device__ explicit Smem(unsigned char* base) {
    b0 = reinterpret_cast<float4*>(base);
    b1 = reinterpret_cast<float4*>(base + Cfg<P, PAIRS>::BUF);
    tw = reinterpret_cast<float2*>(base + 2 * Cfg<P, PAIRS>::BUF);
    postab = reinterpret_cast<unsigned short*>(base + 2 * Cfg<P, PAIRS>::BUF + (size_t)P * sizeof(float2));
  }
  __device__ float4* buf(int x) const { return (x & 1) ? b1 : b0; }
};

// One whole transform of the thread's 12 points.  DIT = false: decimation in frequency (natural in, digit-reversed
// out); DIT = true: its transpose (digit-reversed in, natural out).  `x` counts shared-memory exchanges (buffer x & 1).
template <int P, int PAIRS, int SGN, bool DIT>
__device__ __forceinline__ void transform(C2 (&a)[12], int t, int p, const Smem<P, PAIRS>& sm, int& x) {
  constexpr int NEX = Plan<P>::NEX;
#pragma unroll
  for (int ph = 0; ph <= NEX; ph++) {
    const int pass = DIT ? NEX - ph : ph;
    if (ph > 0) {
      __syncthreads();
      load_pass<P, PAIRS>(pass, a, t, p, sm.buf(x - 1));
    }
    if (DIT) pass_dit<P, SGN>(pass, a, t, sm.tw);
    else pass_dif<P, SGN>(pass, a, t, sm.tw);
    if (ph < NEX) {
      store_pass<P, PAIRS>(pass, a, t, p, sm.buf(x));
      x++;
    }
  }
}

template <int P, int PAIRS>
__device__ __forceinline__ void load_tables(const Smem<P, PAIRS>& sm, const float2* __restrict__ twg, bool want_postab) {
  for (int i = threadIdx.x; i < P; i += Cfg<P, PAIRS>::THREADS) {
    if (i < tw_size<P>()) sm.tw[i] = twg[i];
    if (want_postab) sm.postab[i] = (unsigned short)pos_of_freq<P>(i);
  }
}

// ---- rows forward ----------------------------------------------------------------------------------------------
template <int P, int PAIRS>
__global__ void __launch_bounds__(Cfg<P, PAIRS>::THREADS, Cfg<P, PAIRS>::MINB)
fast_rows_fwd(int nrows, int Ny, int Nyp, int Nx, int Pw, int kxp, const float* __restrict__ rho, float2* __restrict__ A, const float2* __restrict__ twg) {
  using C = Cfg<P, PAIRS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Smem<P, PAIRS> sm(smem_raw);
  const int tid = threadIdx.x, p = tid % PAIRS, t = tid / PAIRS;
  // virtual rows: every instance owns Nyp = Ny rounded up to 4 rows, so that the four rows of a thread (two packed
  // transforms of two real rows each) always belong to ONE instance: an instance's result does not depend on what else
  // is in the batch, bit for bit
  const int r0 = blockIdx.x * C::ROWS + 4 * p;
  const int inst = r0 / Nyp, iy0 = r0 - inst * Nyp;
  load_tables<P, PAIRS>(sm, twg, true);
  pdl_wait();
  pdl_launch_dependents();
  const float* row[4];
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; j++) { ok[j] = r0 + j < nrows && iy0 + j < Ny; row[j] = rho + (size_t)(ok[j] ? inst * Ny + iy0 + j : 0) * Pw; }
  C2 a[12];
#pragma unroll
  for (int e = 0; e < 12; e++) {
    const int ix = t + C::T * e;
    const bool in = ix < Nx;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = (in && ok[j]) ? row[j][ix] : 0.f;
    a[e] = C2{make_float2(v[0], v[2]), make_float2(v[1], v[3])};  // transform .x = rows (0, 1), transform .y = rows (2, 3)
  }
  int x = 0;
  __syncthreads();  // the twiddle / position tables written by load_tables are read by other threads from the first pass on
  transform<P, PAIRS, -1, false>(a, t, p, sm, x);
  store_pass<P, PAIRS>(C::NEX, a, t, p, sm.buf(x));
  x++;
  __syncthreads();
  // Z = FFT(a + i b) -> A_k = (Z_k + conj Z_{P-k}) / 2,  B_k = (Z_k - conj Z_{P-k}) / (2i),  k = 0 .. P/2
  const float4* zb = sm.buf(x - 1);
  float2* out[4];
#pragma unroll
  for (int j = 0; j < 4; j++) out[j] = A + (size_t)(r0 + j) * kxp;
  for (int k = t; k <= P / 2; k += C::T) {
    const C2 zk = unpack(zb[word<PAIRS>(sm.postab[k], p)]), zm = unpack(zb[word<PAIRS>(sm.postab[k ? P - k : 0], p)]);
    const float2 are = pmul(padd(zk.re, zm.re), bc(0.5f)), aim = pmul(psub(zk.im, zm.im), bc(0.5f));
    const float2 bre = pmul(padd(zk.im, zm.im), bc(0.5f)), bim = pmul(psub(zm.re, zk.re), bc(0.5f));
    if (ok[0]) out[0][k] = make_float2(are.x, aim.x);
    if (ok[1]) out[1][k] = make_float2(bre.x, bim.x);
    if (ok[2]) out[2][k] = make_float2(are.y, aim.y);
    if (ok[3]) out[3][k] = make_float2(bre.y, bim.y);
  }
}

// ---- columns: forward, x 3 kernel spectra, inverse --------------------------------------------------------------
template <int P, int PAIRS>
__global__ void __launch_bounds__(Cfg<P, PAIRS>::THREADS, Cfg<P, PAIRS>::MINB)
fast_cols(int Ny, int Nyp, int kxp, int ngroups, const float2* __restrict__ A, const float* __restrict__ hk, float2* __restrict__ G, size_t gstride,
          const float2* __restrict__ twg) {
  using C = Cfg<P, PAIRS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Smem<P, PAIRS> sm(smem_raw);
  const int tid = threadIdx.x, p = tid % PAIRS, t = tid / PAIRS;
  const int g = blockIdx.x, b = blockIdx.y;
  load_tables<P, PAIRS>(sm, twg, false);
  pdl_wait();
  pdl_launch_dependents();
  const size_t col = (size_t)g * C::COLS + 2 * p;               // first of this thread's two columns
  const float2* Ab = A + (size_t)b * Nyp * kxp + col;
  C2 r[12];
#pragma unroll
  for (int e = 0; e < 12; e++) {
    const int iy = t + C::T * e;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy < Ny) v = *reinterpret_cast<const float4*>(Ab + (size_t)iy * kxp);  // (re0, im0, re1, im1)
    r[e] = C2{make_float2(v.x, v.z), make_float2(v.y, v.w)};
  }
  int x = 0;
  __syncthreads();  // tables of load_tables (see fast_rows_fwd): without this barrier a late warp's table entries were read
                    // before they were written -- rare, timing-dependent errors of 1e-7 .. 1e-4 in the force
  transform<P, PAIRS, -1, false>(r, t, p, sm, x);  // r[s] = spectrum at digit-reversed position pass_pos(NEX, t, s)
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    const float* hb = hk + ((size_t)k * ngroups + g) * P * C::COLS + 2 * p;
    C2 w[12];
#pragma unroll
    for (int s = 0; s < 12; s++) {
      const float2 h = *reinterpret_cast<const float2*>(hb + (size_t)pass_pos<P>(C::NEX, t, s) * C::COLS);
      // k = 0: real spectrum, w = r h;  k = 1, 2: purely imaginary spectrum i h, w = (-im h, re h)
      if (k == 0) w[s] = C2{pmul(r[s].re, h), pmul(r[s].im, h)};
      else w[s] = C2{pmul(r[s].im, make_float2(-h.x, -h.y)), pmul(r[s].re, h)};
    }
    transform<P, PAIRS, +1, true>(w, t, p, sm, x);  // natural order: w[e] <-> iy = t + T e
    // G[k][row quad][kx][row mod 4]: the four rows a thread of fast_rows_inv combines sit in one 32-byte sector
    float2* Gk = G + (size_t)k * gstride;
#pragma unroll
    for (int e = 0; e < 12; e++) {
      const int iy = t + C::T * e;
      if (iy < Ny) {
        const size_t r = (size_t)b * Nyp + iy;
        float2* o = Gk + ((r >> 2) * kxp + col) * 4 + (r & 3);
        o[0] = make_float2(w[e].re.x, w[e].im.x);
        o[4] = make_float2(w[e].re.y, w[e].im.y);
      }
    }
  }
}

// ---- rows inverse + division -------------------------------------------------------------------------------------
template <int P, int PAIRS>
__global__ void __launch_bounds__(Cfg<P, PAIRS>::THREADS, Cfg<P, PAIRS>::MINB)
fast_rows_inv(int nrows, int Ny, int Nyp, int Nx, int Pw, int kxp, const float2* __restrict__ G, size_t gstride, float* __restrict__ Fx,
              float* __restrict__ Fy, const float2* __restrict__ twg) {
  using C = Cfg<P, PAIRS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Smem<P, PAIRS> sm(smem_raw);
  const int tid = threadIdx.x, p = tid % PAIRS, t = tid / PAIRS;
  const int r0 = blockIdx.x * C::ROWS + 4 * p;        // virtual row (see fast_rows_fwd): a whole quad of one instance
  const int inst = r0 / Nyp, iy0 = r0 - inst * Nyp;
  const size_t frow = (size_t)inst * Ny + iy0;        // row of Fx / Fy
  load_tables<P, PAIRS>(sm, twg, true);
  pdl_wait();
  pdl_launch_dependents();
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; j++) ok[j] = r0 + j < nrows && iy0 + j < Ny;
  constexpr int NE = 6;  // 2 Nx - 1 <= P = 12 T, so Nx <= 6 T: at most 6 output points per thread and row (ix = t + T i)
  float4 inv[NE];
  int x = 0;
  __syncthreads();  // tables of load_tables
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    const float2* Gk = G + (size_t)k * gstride;
    C2 a[12];
#pragma unroll
    for (int e = 0; e < 12; e++) {
      const int kx = t + C::T * e;
      const bool up = kx > P / 2;           // Hermitian extension: S[kx] = conj(S_a[P - kx]) + i conj(S_b[P - kx])
      const int src = up ? P - kx : kx;
      const float sg = up ? -1.f : 1.f;
      float2 g4[4];
      {  // the block's rows are quad-aligned (r0 = 4 (block PAIRS + p)): two 16-byte loads fetch the four rows at column src
        const float4* q4 = reinterpret_cast<const float4*>(Gk + ((size_t)(r0 >> 2) * kxp + src) * 4);
        const float4 lo = ok[0] ? q4[0] : make_float4(0.f, 0.f, 0.f, 0.f), hi = ok[2] ? q4[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        g4[0] = make_float2(lo.x, lo.y); g4[1] = ok[1] ? make_float2(lo.z, lo.w) : make_float2(0.f, 0.f);
        g4[2] = make_float2(hi.x, hi.y); g4[3] = ok[3] ? make_float2(hi.z, hi.w) : make_float2(0.f, 0.f);
      }
      // rows (0, 1) -> transform .x, rows (2, 3) -> transform .y;  S = A + i B with A, B conjugated when `up`
      a[e] = C2{make_float2(g4[0].x - sg * g4[1].y, g4[2].x - sg * g4[3].y), make_float2(sg * g4[0].y + g4[1].x, sg * g4[2].y + g4[3].x)};
    }
    transform<P, PAIRS, +1, false>(a, t, p, sm, x);  // inverse, digit-reversed output
    store_pass<P, PAIRS>(C::NEX, a, t, p, sm.buf(x));
    x++;
    __syncthreads();
    const float4* ob = sm.buf(x - 1);
    float* F = k == 1 ? Fx : Fy;
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int ix = t + C::T * i;
      if (ix < Nx) {
        const float4 v = ob[word<PAIRS>(sm.postab[ix], p)];  // (re.x, re.y, im.x, im.y) = rows 0, 2, 1, 3
        if (k == 0) {
          inv[i] = make_float4(1.f / v.x, 1.f / v.z, 1.f / v.y, 1.f / v.w);  // rows 0, 1, 2, 3: 1 / (w * rho)
        } else {  // pde.jl:58 force ./ norms
          if (ok[0]) F[(frow + 0) * Pw + ix] = v.x * inv[i].x;
          if (ok[1]) F[(frow + 1) * Pw + ix] = v.z * inv[i].y;
          if (ok[2]) F[(frow + 2) * Pw + ix] = v.y * inv[i].z;
          if (ok[3]) F[(frow + 3) * Pw + ix] = v.w * inv[i].w;
        }
      }
    }
  }
}

// kernel spectra in the fast layout: hk[k][group][pos][COLS] (float) from the Float64 spectra khat[k][kx][slot(ky)] of
// fft_conv.cu (slot = that path's permuted position of frequency ky, given as a table): real part for k = 0, imaginary
// part for k = 1, 2 (the other parts vanish by the symmetry of the kernel images), zero in the pad columns.
template <int P>
__global__ void fast_hk_kernel(int hp1, int cols, int ngroups, const double2* __restrict__ khat, const int* __restrict__ slot_of_freq, float* __restrict__ hk,
                               double* __restrict__ residual) {
  const size_t total = (size_t)3 * ngroups * P * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % cols);
    const int pos = (int)((e / cols) % P);
    const int g = (int)((e / cols / P) % ngroups);
    const int k = (int)(e / cols / P / ngroups);
    const int kx = g * cols + c;
    float v = 0.f;
    if (kx < hp1) {
      const double2 z = khat[((size_t)k * hp1 + kx) * P + slot_of_freq[freq_of_pos<P>(pos)]];
      v = (float)(k == 0 ? z.x : z.y);
      const double drop = fabs(k == 0 ? z.y : z.x);  // the part that symmetry makes vanish: reported, must be rounding only
      if (drop > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(residual), (unsigned long long)__double_as_longlong(drop));
    }
    hk[e] = v;
  }
}

template <int P>
__global__ void fast_tw_kernel(float2* tw) {  // pass-major twiddle table (fft12_core.cuh)
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P) return;
  int num = 0, den = 1;
  if (e < tw_size<P>()) tw_exponent<P>(e, num, den);
  double s, c;
  sincospi(-2.0 * (double)num / (double)den, &s, &c);
  tw[e] = make_float2((float)c, (float)s);
}

}  // namespace

struct smm_fft_fast {
  int P, pairs, Nx, Ny, Pw, B, chunk, kxp, ngroups;
  float2* tw;   // [P] exp(-2 pi i j / P)
  float* hk;    // [3][ngroups][P][2 pairs]
  float2* a;    // [chunk * Ny][kxp]
  float2* g;    // [3][chunk * Ny][kxp]
  double hk_residual;  // largest dropped spectrum component (diagnostic)
};

void smm_fft_fast_destroy(smm_fft_fast* f) {
  if (!f) return;
  void* ptrs[] = {f->tw, f->hk, f->a, f->g};
  for (void* q : ptrs) if (q) cudaFree(q);
  delete f;
}

bool smm_fft_fast_supported(int P) { return P == 96 || P == 192 || P == 384 || P == 768; }

template <int P, int PAIRS> static int fast_set_attr(smm_ctx* ctx) {
  const int sm = (int)Cfg<P, PAIRS>::SMEM;
  SMM_CUDA(ctx, cudaFuncSetAttribute(fast_rows_fwd<P, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  SMM_CUDA(ctx, cudaFuncSetAttribute(fast_cols<P, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  SMM_CUDA(ctx, cudaFuncSetAttribute(fast_rows_inv<P, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  return SMM_OK;
}

template <int P, int PAIRS>
static int fast_launch(smm_ctx* ctx, smm_fft_fast* f, const float* rho, float* Fx, float* Fy) {
  using C = Cfg<P, PAIRS>;
  static const bool pdl = !(getenv("SMM_PDE_PDL") && getenv("SMM_PDE_PDL")[0] == '0');
  const int sms = pdl ? ctx->sm_count : 0;
  for (int b0 = 0; b0 < f->B; b0 += f->chunk) {
    const int bc = std::min(f->chunk, f->B - b0);
    const int Nyp = (f->Ny + 3) & ~3;
    const int nrows = bc * Nyp;                          // virtual rows: Ny rounded up to whole quads per instance
    const size_t off = (size_t)b0 * f->Ny * f->Pw;
    const size_t gstride = (size_t)f->chunk * Nyp * f->kxp;
    const dim3 rgrid((nrows + C::ROWS - 1) / C::ROWS);
    SMM_CUDA(ctx, smm_launch_pdl(sms, fast_rows_fwd<P, PAIRS>, rgrid, dim3(C::THREADS), C::SMEM, ctx->stream, nrows, f->Ny, Nyp, f->Nx, f->Pw,
                                 f->kxp, rho + off, f->a, (const float2*)f->tw));
    SMM_CUDA(ctx, smm_launch_pdl(sms, fast_cols<P, PAIRS>, dim3(f->ngroups, bc), dim3(C::THREADS), C::SMEM, ctx->stream, f->Ny, Nyp, f->kxp,
                                 f->ngroups, (const float2*)f->a, (const float*)f->hk, f->g, gstride, (const float2*)f->tw));
    SMM_CUDA(ctx, smm_launch_pdl(sms, fast_rows_inv<P, PAIRS>, rgrid, dim3(C::THREADS), C::SMEM, ctx->stream, nrows, f->Ny, Nyp, f->Nx, f->Pw,
                                 f->kxp, (const float2*)f->g, gstride, Fx + off, Fy + off, (const float2*)f->tw));
    ctx->launches += 3;
  }
  SMM_CUDA(ctx, cudaGetLastError());
  return SMM_OK;
}

#define FAST_DISPATCH(CALL)                                                                          \
  switch (f->P * 10 + f->pairs) {                                                                    \
    case 964: return CALL(96, 4);   case 968: return CALL(96, 8);                                    \
    case 1924: return CALL(192, 4); case 1928: return CALL(192, 8);                                  \
    case 3844: return CALL(384, 4); case 3848: return CALL(384, 8);                                  \
    case 7684: return CALL(768, 4); case 7688: return CALL(768, 8);                                  \
    default: return smm_fail(ctx, SMM_ERR_UNSUPPORTED, "fast FFT: P=%d pairs=%d", f->P, f->pairs);   \
  }

int smm_fft_fast_conv(smm_ctx* ctx, smm_fft_fast* f, const void* rhosum, void* Fx, void* Fy) {
#define CALL(PP, QQ) fast_launch<PP, QQ>(ctx, f, (const float*)rhosum, (float*)Fx, (float*)Fy)
  FAST_DISPATCH(CALL)
#undef CALL
}

static int fast_attr(smm_ctx* ctx, smm_fft_fast* f) {
#define CALL(PP, QQ) fast_set_attr<PP, QQ>(ctx)
  FAST_DISPATCH(CALL)
#undef CALL
}

// khat64: Float64 kernel spectra [3][hp1][P] in fft_conv.cu's slot order; slot_of_freq_dev: int[P] (device)
int smm_fft_fast_create(smm_ctx* ctx, int P, int Nx, int Ny, int Pw, int B, const void* khat64, const int* slot_of_freq_dev, smm_fft_fast** out) {
  *out = nullptr;
  if (!smm_fft_fast_supported(P)) return smm_fail(ctx, SMM_ERR_UNSUPPORTED, "fast FFT: length %d not supported", P);
  smm_fft_fast* f = new smm_fft_fast();
  memset(f, 0, sizeof *f);
  f->P = P; f->Nx = Nx; f->Ny = Ny; f->Pw = Pw; f->B = B;
  {
    const char* ev = getenv("SMM_FFT_PAIRS");
    f->pairs = (ev && atoi(ev) == 8) ? 8 : 4;
  }
  const int hp1 = P / 2 + 1, cols = 2 * f->pairs;
  f->ngroups = (hp1 + cols - 1) / cols;
  f->kxp = f->ngroups * cols;
  // Instances per launch.  Chunks small enough for the spectra (A and G: 4 x Ny x kxp complex per instance) to stay in L2
  // between the three kernels were measured and lost to whole-batch launches (64 x 353^2: 3.96 vs 3.80 ms per step: the
  // larger grids matter more than the 0.3 GB of extra DRAM traffic); SMM_FFT_CHUNK_MB bounds the scratch when set.
  {
    const size_t per_inst = (size_t)4 * Ny * f->kxp * sizeof(float2);
    f->chunk = B;
    if (const char* ev = getenv("SMM_FFT_CHUNK_MB"))
      f->chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)std::max(1, atoi(ev)) << 20) / per_inst));
  }
  const size_t arow = (size_t)f->chunk * ((Ny + 3) & ~3) * f->kxp;  // virtual rows: Ny rounded up to whole quads per instance
#define FF_CUDA(expr)                                                                                                   \
  do {                                                                                                                  \
    cudaError_t e__ = (expr);                                                                                           \
    if (e__ != cudaSuccess) {                                                                                           \
      smm_fft_fast_destroy(f);                                                                                          \
      cudaGetLastError();                                                                                               \
      return smm_fail(ctx, e__ == cudaErrorMemoryAllocation ? SMM_ERR_NOMEM : SMM_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(e__)); \
    }                                                                                                                   \
  } while (0)
  FF_CUDA(cudaMalloc((void**)&f->tw, (size_t)P * sizeof(float2)));
  FF_CUDA(cudaMalloc((void**)&f->hk, (size_t)3 * f->ngroups * P * cols * sizeof(float)));
  FF_CUDA(cudaMalloc((void**)&f->a, arow * sizeof(float2)));
  FF_CUDA(cudaMalloc((void**)&f->g, 3 * arow * sizeof(float2)));
  FF_CUDA(cudaMemsetAsync(f->a, 0, arow * sizeof(float2), ctx->stream));      // pad columns are read (never written): keep them finite
  FF_CUDA(cudaMemsetAsync(f->g, 0, 3 * arow * sizeof(float2), ctx->stream));
  double* resid = nullptr;
  FF_CUDA(cudaMalloc((void**)&resid, sizeof(double)));
  cudaMemsetAsync(resid, 0, sizeof(double), ctx->stream);
  switch (P) {
    case 96: fast_tw_kernel<96><<<(P + 255) / 256, 256, 0, ctx->stream>>>(f->tw); break;
    case 192: fast_tw_kernel<192><<<(P + 255) / 256, 256, 0, ctx->stream>>>(f->tw); break;
    case 384: fast_tw_kernel<384><<<(P + 255) / 256, 256, 0, ctx->stream>>>(f->tw); break;
    default: fast_tw_kernel<768><<<(P + 255) / 256, 256, 0, ctx->stream>>>(f->tw); break;
  }
  const size_t total = (size_t)3 * f->ngroups * P * cols;
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 8);
  switch (P) {
    case 96: fast_hk_kernel<96><<<grid, 256, 0, ctx->stream>>>(hp1, cols, f->ngroups, (const double2*)khat64, slot_of_freq_dev, f->hk, resid); break;
    case 192: fast_hk_kernel<192><<<grid, 256, 0, ctx->stream>>>(hp1, cols, f->ngroups, (const double2*)khat64, slot_of_freq_dev, f->hk, resid); break;
    case 384: fast_hk_kernel<384><<<grid, 256, 0, ctx->stream>>>(hp1, cols, f->ngroups, (const double2*)khat64, slot_of_freq_dev, f->hk, resid); break;
    default: fast_hk_kernel<768><<<grid, 256, 0, ctx->stream>>>(hp1, cols, f->ngroups, (const double2*)khat64, slot_of_freq_dev, f->hk, resid); break;
  }
  ctx->launches += 2;
  cudaError_t e = cudaMemcpyAsync(&f->hk_residual, resid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(resid);
  FF_CUDA(e);
  FF_CUDA(cudaGetLastError());
#undef FF_CUDA
  int rc = fast_attr(ctx, f);
  if (rc != SMM_OK) { smm_fft_fast_destroy(f); return rc; }
  *out = f;
  return SMM_OK;
}

double smm_fft_fast_residual(const smm_fft_fast* f) { return f ? f->hk_residual : 0.0; }
