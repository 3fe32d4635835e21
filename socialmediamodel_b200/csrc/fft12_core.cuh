// fft12_core.cuh -- register-resident FFT passes for lengths P = 12 T, T in {8, 16, 32, 64} (P = 96, 192, 384, 768:
// the zero-padded lengths 3 * 2^a >= 2N - 1 of every grid the reference and the benchmarks use: 47, 89, 177, 353).
//
// One FFT is worked on by T threads holding 12 complex points each; each thread carries TWO independent transforms
// (two adjacent spectrum columns, or two row pairs) in the halves of packed FP32x2 registers, so every butterfly
// operation is one FADD2 / FMUL2 / FFMA2 with scalar-broadcast twiddles.  Pass structure (decimation in frequency,
// natural order in, digit-reversed order out):
//     pass 1        radix 12 = 3 x 4 over the elements t + T e, e = 0..11, twiddle W_P^(t k1)
//     then          radix-4 passes over quarter distances q = T/4, T/16, .. (three butterflies per thread),
//                   and a final radix-2 pass when log2 T is odd
// The inverse used after the spectral multiplication is the exact transpose (decimation in time, conjugate twiddles,
// digit-reversed in, natural out), so no reordering pass exists anywhere: the kernel spectra are stored in the same
// digit-reversed order.  Between passes the points are exchanged through shared memory in the layout
//     [position i][pair][re0 re1 im0 im1]       (one 16-byte word per thread and point)
// with the pair index fastest across lanes: a quarter warp always reads one contiguous 128-byte row (8 pairs per
// CTA), so every pass is bank-conflict free without padding; with 4 pairs per CTA a quarter warp covers two rows and
// the one-bit swizzle phys(i) = i ^ ((i>>1 ^ i>>2) & 1) keeps the two in different halves of the bank window.
//
// The functions are __host__ __device__ so that tools/fft12_check.cpp can run the very same index arithmetic on the CPU
// (thread by thread, pass by pass) against a direct DFT and count bank conflicts; on the device everything inlines into
// registers (all loops have compile-time bounds).
#pragma once

#if defined(__CUDACC__)
#define F12_HD __host__ __device__ __forceinline__
#else
#define F12_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
#endif

namespace f12 {

// ---- packed arithmetic: lane .x = first transform, lane .y = second transform --------------------------------
#if defined(__CUDA_ARCH__)
F12_HD float2 padd(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
F12_HD float2 psub(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
F12_HD float2 pmul(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
F12_HD float2 pfma(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
#else
F12_HD float2 padd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
F12_HD float2 psub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
F12_HD float2 pmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
F12_HD float2 pfma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
F12_HD float2 bc(float s) { return make_float2(s, s); }  // scalar broadcast (folds into the .F32 operand form)

struct C2 { float2 re, im; };  // one complex point of each of the two transforms
F12_HD C2 cadd(C2 a, C2 b) { return C2{padd(a.re, b.re), padd(a.im, b.im)}; }
F12_HD C2 csub(C2 a, C2 b) { return C2{psub(a.re, b.re), psub(a.im, b.im)}; }
// a * (wr + i wi), the twiddle shared by both transforms
F12_HD C2 cmul_tw(C2 a, float wr, float wi) {
  return C2{pfma(a.re, bc(wr), pmul(a.im, bc(-wi))), pfma(a.re, bc(wi), pmul(a.im, bc(wr)))};
}
// sign-aware: SGN = -1 forward (multiply by w), +1 inverse (multiply by conj(w)), w = wr + i wi the FORWARD twiddle
template <int SGN> F12_HD C2 cmul_s(C2 a, float wr, float wi) { return SGN < 0 ? cmul_tw(a, wr, wi) : cmul_tw(a, wr, -wi); }
// multiplication by -i (forward) / +i (inverse): SGN * i
template <int SGN> F12_HD C2 mul_si(C2 a) {
  // (re + i im) * (s i) = -s im + i s re
  if (SGN < 0) return C2{a.im, psub(make_float2(0.f, 0.f), a.re)};  // * (-i) = (im, -re)
  return C2{psub(make_float2(0.f, 0.f), a.im), a.re};               // * (+i) = (-im, re)
}

// ---- small DFTs on registers; SGN = -1: exp(-2 pi i ..), +1: exp(+2 pi i ..) ----------------------------------
template <int SGN> F12_HD void dft4(C2& a0, C2& a1, C2& a2, C2& a3) {
  const C2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  // y1 = d02 + (SGN i) d13, y3 = d02 - (SGN i) d13;  (SGN i)(x + i y) = -SGN y + i SGN x
  C2 y1, y3;
  if (SGN < 0) {  // -i d13 = (d13.im, -d13.re)
    y1 = C2{padd(d02.re, d13.im), psub(d02.im, d13.re)};
    y3 = C2{psub(d02.re, d13.im), padd(d02.im, d13.re)};
  } else {
    y1 = C2{psub(d02.re, d13.im), padd(d02.im, d13.re)};
    y3 = C2{padd(d02.re, d13.im), psub(d02.im, d13.re)};
  }
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = y1;
  a3 = y3;
}
template <int SGN> F12_HD void dft3(C2& a0, C2& a1, C2& a2) {
  // w = exp(SGN 2 pi i / 3) = -1/2 + SGN i sqrt(3)/2;  y1 = a0 + w a1 + w^2 a2, y2 = a0 + w^2 a1 + w a2
  const float h = 0.86602540378443864676f;
  const C2 s = cadd(a1, a2), d = csub(a1, a2);
  const C2 t = C2{pfma(s.re, bc(-0.5f), a0.re), pfma(s.im, bc(-0.5f), a0.im)};
  // (SGN i h) d = (-SGN h d.im, SGN h d.re)
  const float c = SGN < 0 ? h : -h;  // u = (c d.im, -c d.re)
  const C2 y1 = C2{pfma(d.im, bc(c), t.re), pfma(d.re, bc(-c), t.im)};
  const C2 y2 = C2{pfma(d.im, bc(-c), t.re), pfma(d.re, bc(c), t.im)};
  a0 = cadd(a0, s);
  a1 = y1;
  a2 = y2;
}

// W_12^k (forward, exp(-2 pi i k / 12)) as constants
F12_HD void w12(int k, float& wr, float& wi) {
  const float c30 = 0.86602540378443864676f;
  switch (k) {
    case 0: wr = 1.f; wi = 0.f; break;
    case 1: wr = c30; wi = -0.5f; break;
    case 2: wr = 0.5f; wi = -c30; break;
    case 3: wr = 0.f; wi = -1.f; break;
    case 4: wr = -0.5f; wi = -c30; break;
    default: wr = -1.f; wi = 0.f; break;  // 6
  }
}

// 12-point DFT, decimation in frequency form: input a[e] (e = 4 e3 + e4), output in place a[k1] (k1 = k3 + 3 k4):
// 3-point DFTs over e3, twiddle W_12^(e4 k3), 4-point DFTs over e4.
template <int SGN> F12_HD void dft12_dif(C2 (&a)[12]) {
#pragma unroll
  for (int e4 = 0; e4 < 4; e4++) dft3<SGN>(a[e4], a[4 + e4], a[8 + e4]);  // now a[4 k3 + e4] = b[k3][e4]
  // twiddles W_12^(e4 k3): k3 = 1: e4 = 1,2,3 -> W^1, W^2, W^3;  k3 = 2: e4 = 1,2,3 -> W^2, W^4, W^6
  {
    float wr, wi;
    w12(1, wr, wi); a[4 + 1] = cmul_s<SGN>(a[4 + 1], wr, wi);
    w12(2, wr, wi); a[4 + 2] = cmul_s<SGN>(a[4 + 2], wr, wi);
    a[4 + 3] = mul_si<SGN>(a[4 + 3]);  // W^3 = -i (forward)
    w12(2, wr, wi); a[8 + 1] = cmul_s<SGN>(a[8 + 1], wr, wi);
    w12(4, wr, wi); a[8 + 2] = cmul_s<SGN>(a[8 + 2], wr, wi);
    a[8 + 3] = C2{psub(make_float2(0.f, 0.f), a[8 + 3].re), psub(make_float2(0.f, 0.f), a[8 + 3].im)};  // W^6 = -1
  }
  C2 y[12];
#pragma unroll
  for (int k3 = 0; k3 < 3; k3++) {
    C2 b0 = a[4 * k3], b1 = a[4 * k3 + 1], b2 = a[4 * k3 + 2], b3 = a[4 * k3 + 3];
    dft4<SGN>(b0, b1, b2, b3);
    y[k3] = b0; y[k3 + 3] = b1; y[k3 + 6] = b2; y[k3 + 9] = b3;
  }
#pragma unroll
  for (int k = 0; k < 12; k++) a[k] = y[k];
}
// transpose of dft12_dif<-SGN'>: input y[k1] (k1 = k3 + 3 k4), output a[e] (e = 4 e3 + e4), kernel exp(SGN 2 pi i e k1 / 12)
template <int SGN> F12_HD void dft12_dit(C2 (&y)[12]) {
  C2 c[12];  // c[4 k3 + e4]
#pragma unroll
  for (int k3 = 0; k3 < 3; k3++) {
    C2 b0 = y[k3], b1 = y[k3 + 3], b2 = y[k3 + 6], b3 = y[k3 + 9];
    dft4<SGN>(b0, b1, b2, b3);
    c[4 * k3] = b0; c[4 * k3 + 1] = b1; c[4 * k3 + 2] = b2; c[4 * k3 + 3] = b3;
  }
  {
    float wr, wi;
    w12(1, wr, wi); c[4 + 1] = cmul_s<SGN>(c[4 + 1], wr, wi);
    w12(2, wr, wi); c[4 + 2] = cmul_s<SGN>(c[4 + 2], wr, wi);
    c[4 + 3] = mul_si<SGN>(c[4 + 3]);
    w12(2, wr, wi); c[8 + 1] = cmul_s<SGN>(c[8 + 1], wr, wi);
    w12(4, wr, wi); c[8 + 2] = cmul_s<SGN>(c[8 + 2], wr, wi);
    c[8 + 3] = C2{psub(make_float2(0.f, 0.f), c[8 + 3].re), psub(make_float2(0.f, 0.f), c[8 + 3].im)};
  }
#pragma unroll
  for (int e4 = 0; e4 < 4; e4++) dft3<SGN>(c[e4], c[4 + e4], c[8 + e4]);  // c[4 e3 + e4] = a[4 e3 + e4]
#pragma unroll
  for (int k = 0; k < 12; k++) y[k] = c[k];
}

// ---- plan constants ------------------------------------------------------------------------------------------
template <int P> struct Plan {
  static_assert(P == 96 || P == 192 || P == 384 || P == 768, "P = 12 T with T in {8, 16, 32, 64}");
  static constexpr int T = P / 12;
  static constexpr int LOG2T = T == 8 ? 3 : T == 16 ? 4 : T == 32 ? 5 : 6;
  static constexpr int NR4 = LOG2T / 2;       // radix-4 passes after the radix-12 pass
  static constexpr bool R2 = (LOG2T & 1) != 0;  // final radix-2 pass
  static constexpr int NEX = NR4 + (R2 ? 1 : 0);  // shared-memory exchanges of one transform
};

// frequency of a digit-reversed position and its inverse (host and device; used for the tables only)
template <int P> F12_HD int freq_of_pos(int pos) {
  constexpr int T = Plan<P>::T;
  const int k1 = pos / T;
  int rest = pos - k1 * T, len = T, mul = 12, f = k1;
  while (len >= 4) { const int q = len >> 2, d = rest / q; rest -= d * q; f += mul * d; mul *= 4; len = q; }
  if (len == 2) { f += mul * rest; }
  return f;
}
template <int P> F12_HD int pos_of_freq(int k) {
  constexpr int T = Plan<P>::T;
  const int k1 = k % 12;
  int k2 = k / 12, len = T, pos = k1 * T;
  while (len >= 4) { const int q = len >> 2; pos += (k2 & 3) * q; k2 >>= 2; len = q; }
  if (len == 2) pos += k2 & 1;
  return pos;
}

// ---- shared-memory addressing --------------------------------------------------------------------------------
// word index (in float4 units) of position i of pair p
template <int PAIRS> F12_HD int swz(int i) { return PAIRS == 8 ? i : (i ^ (((i >> 1) ^ (i >> 2)) & 1)); }
template <int PAIRS> F12_HD int word(int i, int p) { return swz<PAIRS>(i) * PAIRS + p; }

F12_HD float4 pack(C2 v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
F12_HD C2 unpack(float4 v) { return C2{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }

// Position of slot s of thread t in each pass.  Slot numbering: radix-12 pass: s = e (or k1); radix-4 pass: s = 4 u + m
// (butterfly u = 0..2, leg m); radix-2 pass: s = 2 u + m (u = 0..5).
template <int P> F12_HD int pos_r12(int t, int s) { return t + Plan<P>::T * s; }
template <int P> F12_HD int pos_r4(int t, int q, int s) {  // q = quarter distance (power of two, 1 <= q <= T/4)
  // butterfly id = t + T u: leg j = id mod q = t mod q (T is a multiple of q), block = id / q = t / q + (T / q) u;
  // position = block * 4q + j + q m, written as (thread part) + (compile-time part)
  const int u = s >> 2, m = s & 3;
  return (t / q) * 4 * q + (t & (q - 1)) + 4 * Plan<P>::T * u + q * m;
}
template <int P> F12_HD int pos_r2(int t, int s) { return 2 * (t + Plan<P>::T * (s >> 1)) + (s & 1); }

// position of slot s after the LAST forward pass (= before the first inverse pass)
template <int P> F12_HD int pos_last(int t, int s) { return Plan<P>::R2 ? pos_r2<P>(t, s) : pos_r4<P>(t, 1, s); }

// ---- twiddle table, pass-major: every pass reads consecutive entries for consecutive threads --------------------
//   [0, 11 T)                 radix-12 pass:  entry (k1 - 1) T + t        = exp(-2 pi i t k1 / P),  k1 = 1..11
//   then per radix-4 pass with quarter distance q > 1 (in pass order):  entry off + (k - 1) q + j = exp(-2 pi i j k / (4 q)), k = 1..3
template <int P> F12_HD constexpr int tw_r4_offset(int pass) {  // pass = 1..NR4
  int off = 11 * Plan<P>::T;
  for (int ps = 1; ps < pass; ps++) { const int q = Plan<P>::T >> (2 * ps); if (q > 1) off += 3 * q; }
  return off;
}
template <int P> F12_HD constexpr int tw_size() { return tw_r4_offset<P>(Plan<P>::NR4 + 1); }
// exponent of entry e as a fraction num / den of a full turn (for the table generators on host and device)
template <int P> F12_HD void tw_exponent(int e, int& num, int& den) {
  constexpr int T = Plan<P>::T;
  if (e < 11 * T) { const int k1 = e / T + 1, t = e - (k1 - 1) * T; num = t * k1; den = P; return; }
  for (int ps = 1; ps <= Plan<P>::NR4; ps++) {
    const int q = T >> (2 * ps);
    if (q <= 1) continue;
    const int off = tw_r4_offset<P>(ps);
    if (e < off + 3 * q) { const int k = (e - off) / q + 1, j = (e - off) - (k - 1) * q; num = j * k; den = 4 * q; return; }
  }
  num = 0; den = 1;
}

// ---- passes on registers -------------------------------------------------------------------------------------
// SGN as above; tw = the pass-major table.
template <int P, int SGN> F12_HD void twiddle_r12_out(C2 (&a)[12], int t, const float2* tw) {  // a[k1] *= W^(SGN: conj)(t k1)
#pragma unroll
  for (int k1 = 1; k1 < 12; k1++) {
    const float2 w = tw[(k1 - 1) * Plan<P>::T + t];
    a[k1] = cmul_s<SGN>(a[k1], w.x, w.y);
  }
}
// DIF radix-4 pass on the thread's three butterflies: y_k = (sum_m a_m W4^(mk)) * W_{4q}^(j k)
template <int P, int SGN> F12_HD void r4_dif(C2 (&a)[12], int t, int pass, const float2* tw) {
  const int q = Plan<P>::T >> (2 * pass);
  const int j = t & (q - 1);  // T is a multiple of q
  float2 w1 = make_float2(1.f, 0.f), w2 = w1, w3 = w1;
  if (q > 1) { const float2* tq = tw + tw_r4_offset<P>(pass) + j; w1 = tq[0]; w2 = tq[q]; w3 = tq[2 * q]; }
#pragma unroll
  for (int u = 0; u < 3; u++) {
    dft4<SGN>(a[4 * u], a[4 * u + 1], a[4 * u + 2], a[4 * u + 3]);
    if (q > 1) {
      a[4 * u + 1] = cmul_s<SGN>(a[4 * u + 1], w1.x, w1.y);
      a[4 * u + 2] = cmul_s<SGN>(a[4 * u + 2], w2.x, w2.y);
      a[4 * u + 3] = cmul_s<SGN>(a[4 * u + 3], w3.x, w3.y);
    }
  }
}
// DIT radix-4 pass (transpose of r4_dif with the opposite sign): twiddle first, then the butterfly
template <int P, int SGN> F12_HD void r4_dit(C2 (&a)[12], int t, int pass, const float2* tw) {
  const int q = Plan<P>::T >> (2 * pass);
  const int j = t & (q - 1);
  float2 w1 = make_float2(1.f, 0.f), w2 = w1, w3 = w1;
  if (q > 1) { const float2* tq = tw + tw_r4_offset<P>(pass) + j; w1 = tq[0]; w2 = tq[q]; w3 = tq[2 * q]; }
#pragma unroll
  for (int u = 0; u < 3; u++) {
    if (q > 1) {
      a[4 * u + 1] = cmul_s<SGN>(a[4 * u + 1], w1.x, w1.y);
      a[4 * u + 2] = cmul_s<SGN>(a[4 * u + 2], w2.x, w2.y);
      a[4 * u + 3] = cmul_s<SGN>(a[4 * u + 3], w3.x, w3.y);
    }
    dft4<SGN>(a[4 * u], a[4 * u + 1], a[4 * u + 2], a[4 * u + 3]);
  }
}
F12_HD void r2_pass(C2 (&a)[12]) {
#pragma unroll
  for (int u = 0; u < 6; u++) {
    const C2 s = cadd(a[2 * u], a[2 * u + 1]), d = csub(a[2 * u], a[2 * u + 1]);
    a[2 * u] = s;
    a[2 * u + 1] = d;
  }
}

// ---- whole transforms as sequences of passes; pass 0 = radix 12, 1..NR4 = radix 4 (q = T >> 2 pass), NR4+1 = radix 2 ----
template <int P> F12_HD int pass_pos(int pass, int t, int s) {
  if (pass == 0) return pos_r12<P>(t, s);
  if (pass <= Plan<P>::NR4) return pos_r4<P>(t, Plan<P>::T >> (2 * pass), s);
  return pos_r2<P>(t, s);
}
template <int P, int SGN> F12_HD void pass_dif(int pass, C2 (&a)[12], int t, const float2* tw) {
  if (pass == 0) { dft12_dif<SGN>(a); twiddle_r12_out<P, SGN>(a, t, tw); }
  else if (pass <= Plan<P>::NR4) r4_dif<P, SGN>(a, t, pass, tw);
  else r2_pass(a);
}
template <int P, int SGN> F12_HD void pass_dit(int pass, C2 (&a)[12], int t, const float2* tw) {
  if (pass == 0) { twiddle_r12_out<P, SGN>(a, t, tw); dft12_dit<SGN>(a); }
  else if (pass <= Plan<P>::NR4) r4_dit<P, SGN>(a, t, pass, tw);
  else r2_pass(a);
}
// Shared-memory word of slot s in a pass = (one of two per-thread bases) + (compile-time offset).  The swizzle bit of a
// position depends either on thread bits only or on slot bits only in every pass (pos = thread part + slot part, and
// the two never carry into bits 0..2 of each other), so the XOR of swz() reduces to a choice between the thread part
// with bit 0 as is / flipped: no per-access address arithmetic (the generic form cost 30 registers of spills).
// tools/fft12_check.cpp checks slot_word == word(pass_pos) for every thread, slot and pass.
struct PassAddr { int a, b; };
template <int P, int PAIRS> F12_HD PassAddr pass_bases(int pass, int t, int p) {
  constexpr int T = Plan<P>::T, NR4 = Plan<P>::NR4;
  int tp, fa = 0, fb = 0;
  if (pass == 0) { tp = t; fa = fb = ((t >> 1) ^ (t >> 2)) & 1; }
  else if (pass <= NR4) {
    const int q = T >> (2 * pass);
    tp = (t / q) * 4 * q + (t & (q - 1));
    if (q >= 8) fa = fb = ((t >> 1) ^ (t >> 2)) & 1;
    else if (q == 4) { fa = (t >> 1) & 1; fb = fa ^ 1; }
    else if (q == 2) { fa = 0; fb = 1; }
    else { fa = t & 1; fb = fa ^ 1; }      // q == 1: tp = 4 t
  } else { tp = 2 * t; fa = (t ^ (t >> 1)) & 1; fb = fa ^ 1; }
  if (PAIRS == 8) fa = fb = 0;
  return PassAddr{(tp ^ fa) * PAIRS + p, (tp ^ fb) * PAIRS + p};
}
template <int P, int PAIRS> F12_HD int slot_word(int pass, int s, PassAddr ad) {
  constexpr int T = Plan<P>::T, NR4 = Plan<P>::NR4;
  if (pass == 0) return ad.a + T * s * PAIRS;
  if (pass <= NR4) {
    const int q = T >> (2 * pass), u = s >> 2, m = s & 3, cp = 4 * T * u + q * m;
    if (PAIRS == 8 || q >= 8) return ad.a + cp * PAIRS;
    if (q == 4) return ((m & 1) ? ad.b : ad.a) + cp * PAIRS;
    if (q == 2) return ((((m & 1) ^ (m >> 1)) & 1) ? ad.b : ad.a) + cp * PAIRS;
    // q == 1: positions tp + 4 T u + (m ^ s), s = (m >> 1) ^ (t & 1): m = 0 -> A, 1 -> B, 2 -> B + 2, 3 -> A + 2
    return ((m == 0 || m == 3) ? ad.a : ad.b) + (4 * T * u + (m & 2)) * PAIRS;
  }
  const int u = s >> 1, m = s & 1;
  if (PAIRS == 8) return ad.a + (2 * T * u + m) * PAIRS;
  return (m ? ad.b : ad.a) + 2 * T * u * PAIRS;
}
template <int P, int PAIRS> F12_HD void store_pass(int pass, const C2 (&a)[12], int t, int p, float4* buf) {
  const PassAddr ad = pass_bases<P, PAIRS>(pass, t, p);
#pragma unroll
  for (int s = 0; s < 12; s++) buf[slot_word<P, PAIRS>(pass, s, ad)] = pack(a[s]);
}
template <int P, int PAIRS> F12_HD void load_pass(int pass, C2 (&a)[12], int t, int p, const float4* buf) {
  const PassAddr ad = pass_bases<P, PAIRS>(pass, t, p);
#pragma unroll
  for (int s = 0; s < 12; s++) a[s] = unpack(buf[slot_word<P, PAIRS>(pass, s, ad)]);
}

}  // namespace f12
