// pde_cluster.cu -- K5..K9 for ONE small instance (config C3: PDEsolve on the default 89 x 89 grid): whole RK steps
// inside one thread-block cluster of 16 CTAs, the state resident in the cluster's shared memory.
//
// The kernel chain of pde.cu / fft_conv.cu spends a single 89 x 89 instance on 16 dependent launches per RK4 step
// (4.5-12 us each: 88 us per step for 0.25 MB of state).  Here the 16 CTAs of a cluster own 6 grid rows (and 7 spectrum
// columns) each; a stage is
//   P1  row transforms of the density sum (two real rows per complex FFT), half spectra scattered to the CTAs that own
//       the spectrum columns                                                        -- distributed shared memory stores
//   P2  column transforms: forward, times the three kernel spectra (resident in shared memory), three inverse
//       transforms at once; the kept rows go back to the CTAs that own the grid rows -- distributed shared memory stores
//   P3  inverse row transforms, F = (K rho) / (W rho) (pde.jl:49-59); halo rows of the state from the neighbour CTAs
//   P4  advection + reaction + diffusion + boundary rows (pde.jl:204-236), RK update, density sum and moments of the
//       new stage state; the 24 partial moments go to every CTA of the cluster
//   P5  every CTA adds the 16 partial moment sets in the same order and advances the influencer / media ODEs
//       (pde.jl:239-265) redundantly in Float64: no broadcast, no last-CTA ticket
// with one hardware cluster barrier (barrier.cluster) after P1..P4 instead of a kernel boundary, and no HBM or L2
// traffic at all between the first and the last step of a call.  Arithmetic per grid point = pde_stage_kernel's.
#include <cooperative_groups.h>

#include "fft_generic.cuh"
#include "pde_types.cuh"
#include "smm_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int CS = 16;    // CTAs per cluster (non-portable size: opt-in attribute)
constexpr int NTC = 576;  // threads per CTA: 6 rows x 89 points = 534 grid points per CTA, one thread each

struct Lay {  // shared-memory layout (byte offsets); the same function runs on the host (size) and on the device
  int NP, RPC, CPC, Nx, Ny, P, cs, hp1, NXP, npair;
  size_t U[2], un, acc, Fx, Fy, rho, RW, R, Wk, kh, Arecv, Grecv, tw, mom, part, sm, end;
};
__host__ __device__ inline Lay make_lay(int L, int Nx, int Ny, int P, int RPC, int CPC) {
  Lay l;
  l.NP = 2 * L; l.RPC = RPC; l.CPC = CPC; l.Nx = Nx; l.Ny = Ny; l.P = P; l.cs = P + 1; l.hp1 = P / 2 + 1; l.NXP = Nx + 2;
  l.npair = RPC / 2;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
  for (int b = 0; b < 2; b++) l.U[b] = take(sizeof(float) * l.NP * (RPC + 2) * l.NXP);  // stage state, one-cell halo
  l.un = take(sizeof(float) * l.NP * RPC * Nx);       // state at the start of the step
  l.acc = take(sizeof(float) * l.NP * RPC * Nx);      // RK accumulator
  l.Fx = take(sizeof(float) * RPC * l.NXP);           // x halo
  l.Fy = take(sizeof(float) * (RPC + 2) * Nx);        // y halo
  l.rho = take(sizeof(float) * RPC * Nx);             // density sum of the stage state
  l.R = take(sizeof(float2) * CPC * l.cs);            // column transforms: spectrum of the density sum
  {                                                   // row transforms (P1: npair, P3: 3 npair columns) share their memory with
    const size_t rw = sizeof(float2) * 3 * (l.npair + 1) * l.cs, wk = sizeof(float2) * 3 * CPC * l.cs;  // the three products /
    l.Wk = l.RW = take(rw > wk ? rw : wk);            // inverse column transforms (P2): cluster barriers separate the uses
  }
  l.kh = take(sizeof(float2) * 3 * CPC * P);          // kernel spectra of this CTA's columns
  l.Arecv = take(sizeof(float2) * CPC * Ny);          // written by the row owners (P1)
  l.Grecv = take(sizeof(float2) * 3 * (RPC + 2) * l.hp1);  // written by the column owners (P2): own rows 1..RPC, halo rows 0, nr+1
  l.tw = take(sizeof(float2) * P);
  l.mom = take(sizeof(double) * CS * l.NP * 3);       // partial moments of every CTA (written by all, P4)
  l.part = take(sizeof(double) * l.NP * 3);
  l.sm = take(sizeof(double) * l.NP * 3);
  l.end = o;
  return l;
}

// -DSMM_CLUSTER_DEBUG: globaltimer stamps at the phase boundaries of one stage (tools/scratch/cluster_prof.py reads them)
#ifdef SMM_CLUSTER_DEBUG
__device__ unsigned long long g_dbg[64];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DBG(i) do { if (rank == 3 && tid == 0 && step == 1 && s == 1) g_dbg[i] = gtime(); } while (0)
#else
#define DBG(i) do { } while (0)
#endif
__device__ __forceinline__ double r_fun_c(double x) { return x > 0.1 ? x : 0.1; }  // pde.jl:26-32: r(NaN) = 0.1

// LT influencers; N x N grid; FFT length P_ = 3 * 2^a; RPC_ grid rows and CPC_ spectrum columns per CTA -- compile-time, so
// that every index of the kernel is a shift, a multiplication or a thread constant (the first version, with run-time
// shapes, spent half of its 74 000 warp instructions per CTA and stage on integer divisions).
template <int LT, int N, int P_, int RPC_, int CPC_>
__global__ void __launch_bounds__(NTC, 1) pde_cluster_kernel(const PdeClusterArgs A) {
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int L = LT, NP = 2 * LT, NV = NP * 3;
  constexpr int Nx = N, Ny = N, P = P_, RPC = RPC_, CPC = CPC_;
  constexpr int M_ = P_ / 3, A_ = M_ == 64 ? 6 : (M_ == 32 ? 5 : -1);
  static_assert(A_ > 0 && NTC % P == 0 && RPC * Nx <= NTC && (RPC % 2) == 0, "unsupported shape");
  constexpr FftDims d = {P, M_, A_, 3};
  constexpr int cs = P + 1, hp1 = P / 2 + 1, NXP = Nx + 2, nP = NTC / P, nH = NTC / hp1, nR = NTC / Nx;
  constexpr int UPS = (RPC + 2) * NXP;  // stride between influencers in the state tiles (float2 units)
  const int rank = (int)cluster.block_rank(), b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NTC / 32;
  const Lay l = make_lay(L, Nx, Ny, P, RPC, CPC);
  extern __shared__ __align__(16) unsigned char smem[];
  // state tiles: the two media planes of an influencer interleaved, U[j][row][x] = (u_{j,medium 0}, u_{j,medium 1}): the stencil
  // treats them as one packed FP32x2 value (same formulas, per-medium constants in the two halves)
  float2* const U0 = reinterpret_cast<float2*>(smem + l.U[0]);
  float2* const U1 = reinterpret_cast<float2*>(smem + l.U[1]);
  float2* un = reinterpret_cast<float2*>(smem + l.un);    // [j][row][x] pairs, like the tiles
  float2* acc = reinterpret_cast<float2*>(smem + l.acc);
  float* sFx = reinterpret_cast<float*>(smem + l.Fx);
  float* sFy = reinterpret_cast<float*>(smem + l.Fy);
  float* rho = reinterpret_cast<float*>(smem + l.rho);
  float2* RW = reinterpret_cast<float2*>(smem + l.RW);
  float2* R = reinterpret_cast<float2*>(smem + l.R);
  float2* Wk = reinterpret_cast<float2*>(smem + l.Wk);
  float2* kh = reinterpret_cast<float2*>(smem + l.kh);
  float2* Arecv = reinterpret_cast<float2*>(smem + l.Arecv);
  float2* Grecv = reinterpret_cast<float2*>(smem + l.Grecv);
  float2* tw = reinterpret_cast<float2*>(smem + l.tw);
  double* mom_all = reinterpret_cast<double*>(smem + l.mom);
  double* part = reinterpret_cast<double*>(smem + l.part);
  double* sm = reinterpret_cast<double*>(smem + l.sm);
  // ODE unknowns and stage coefficients (Float64), replicated in every CTA
  __shared__ double c_y[LMAX][2], c_z[2][2], c_dy[LMAX][2], c_dz[2][2], c_R[2][LMAX];
  __shared__ double o_yn[LMAX][2], o_zn[2][2], o_ya[LMAX][2], o_za[2][2];
  __shared__ float s_yx[LMAX], s_yy[LMAX];
  __shared__ float2 s_R2[LMAX], s_bx2[LMAX], s_by2[LMAX];  // per influencer, the two media in the halves: rates, b y_j + c z_i

  const int r0 = min(rank * RPC, Ny), r1 = min(r0 + RPC, Ny), nr = r1 - r0;    // grid rows of this CTA
  const int k0 = min(rank * CPC, hp1), k1 = min(k0 + CPC, hp1), nc = k1 - k0;  // spectrum columns of this CTA
  const int np = (nr + 1) >> 1;                                                // row pairs
  const size_t plane = (size_t)A.Pw * Ny;
  float* ug = A.u + (size_t)b * NP * plane;
  float* rhog = A.rhosum + (size_t)b * plane;
  PdeCoef& cfg = static_cast<PdeCoef*>(A.coef)[b];
  PdeOde& odg = static_cast<PdeOde*>(A.ode)[b];
  const PdeInst& in = static_cast<const PdeInst*>(A.inst)[b];

  // ---- thread constants: every loop below runs over (this thread's element, a short strided range) ------------------------
  const int tP = tid % P, cP = tid / P, swzP = swz(tP);          // element of an FFT column, column group
  const int tx = tid % Nx, tr = tid / Nx, swzX = swz(tx);        // grid column, row slot (tr < nR)
  const bool pt = tr < nr;                                       // this thread owns grid point (r0 + tr, tx)
  const int tk = tid % hp1, ck = tid / hp1;                      // spectrum column kx = tk, row-pair slot (ck < nH)
  const int tk_pk = swz(pos_of_freq(tk, d)), tk_pm = swz(pos_of_freq(tk ? P - tk : 0, d));
  const int tk_owner = tk / CPC, tk_kc = tk - tk_owner * CPC;    // CTA that owns spectrum column tk
  constexpr int nG = NTC / CPC;                                  // P2 scatter: thread = (column gc fastest, grid row gr, gr + nG, ..)
  const int gc = tid % CPC, gr = tid / CPC;                      //   consecutive lanes write consecutive words of the owner's inbox
  const int fp_k = freq_of_pos(tP, d);                           // P3: spectrum index that belongs at FFT position tP
  const bool fp_up = fp_k > P / 2;
  const int fp_src = fp_up ? P - fp_k : fp_k;

  // ---- load: tables, this CTA's rows of the state, the ODE unknowns ----------------------------------------------------
  for (size_t e = tid; e < (l.un - l.U[0]) / 8; e += NTC) U0[e] = make_float2(0.f, 0.f);  // both state buffers incl. halos
  for (int e = tid; e < RPC * NXP; e += NTC) sFx[e] = 0.f;
  for (int e = tid; e < (RPC + 2) * Nx; e += NTC) sFy[e] = 0.f;
  for (int e = tid; e < P; e += NTC) tw[e] = static_cast<const float2*>(A.tw)[e];
  for (int k3 = 0; k3 < 3; k3++)
    for (int c = cP; c < nc; c += nP) kh[(k3 * CPC + c) * P + tP] = static_cast<const float2*>(A.khat)[((size_t)k3 * hp1 + k0 + c) * P + tP];
  __syncthreads();
  if (pt) {
    for (int j = 0; j < L; j++) {
      const size_t o = (size_t)(r0 + tr) * A.Pw + tx;
      U0[j * UPS + (tr + 1) * NXP + tx + 1] = make_float2(ug[(size_t)(2 * j) * plane + o], ug[(size_t)(2 * j + 1) * plane + o]);
    }
    rho[tr * Nx + tx] = rhog[(size_t)(r0 + tr) * A.Pw + tx];
  }
  if (tid < L) {
    for (int dd = 0; dd < 2; dd++) {
      c_y[tid][dd] = cfg.y[tid][dd]; c_dy[tid][dd] = cfg.dy[tid][dd]; o_yn[tid][dd] = odg.yn[tid][dd]; o_ya[tid][dd] = 0.0;
      c_R[dd][tid] = cfg.R[dd][tid];
    }
  }
  if (tid < 2) {
    for (int dd = 0; dd < 2; dd++) { c_z[tid][dd] = cfg.z[tid][dd]; c_dz[tid][dd] = cfg.dz[tid][dd]; o_zn[tid][dd] = odg.zn[tid][dd]; o_za[tid][dd] = 0.0; }
  }
  __syncthreads();
  const float pb = (float)in.b, pc = (float)in.c;
  auto refresh_f32 = [&]() {
    if (tid < L) {
      const float yx = (float)c_y[tid][0], yy = (float)c_y[tid][1];
      s_yx[tid] = yx; s_yy[tid] = yy;
      s_R2[tid] = make_float2((float)c_R[0][tid], (float)c_R[1][tid]);
      s_bx2[tid] = make_float2(pb * yx + pc * (float)c_z[0][0], pb * yx + pc * (float)c_z[1][0]);
      s_by2[tid] = make_float2(pb * yy + pc * (float)c_z[0][1], pb * yy + pc * (float)c_z[1][1]);
    }
  };
  refresh_f32();
  cluster.sync();  // every CTA of the cluster is resident and initialised before the first remote store

  const float pa = (float)in.a, D = (float)in.D;
  const float h = (float)A.dx, ih = (float)(1.0 / A.dx), ih2 = (float)(1.0 / (A.dx * A.dx));
  const float Ddx = (float)(in.D / A.dx);  // pde.jl:231-234 (Q12); the reference builds the boundary rows from dx for both axes
  const float bc = pb + pc;
  const double dV = A.dx * A.dx, dt = A.dt;
  // this thread's grid point
  const int gy = r0 + tr;
  const float X = (float)(A.gx0 + tx * A.dx), Y = (float)(A.gy0 + gy * A.dx);
  const bool x_lo = tx == 0, x_hi = tx == Nx - 1, y_lo = gy == 0, y_hi = gy == Ny - 1;
  // remote views of the neighbours' tiles (rows above / below this CTA's)
  const bool has_up = nr > 0 && r0 > 0, has_dn = nr > 0 && r1 < Ny;
  const float2* const upU0 = has_up ? cluster.map_shared_rank(U0, rank - 1) : nullptr;
  const float2* const upU1 = has_up ? cluster.map_shared_rank(U1, rank - 1) : nullptr;
  const float2* const dnU0 = has_dn ? cluster.map_shared_rank(U0, rank + 1) : nullptr;
  const float2* const dnU1 = has_dn ? cluster.map_shared_rank(U1, rank + 1) : nullptr;
  float2* A_dst = cluster.map_shared_rank(Arecv, tk_owner) + tk_kc * Ny + r0;          // P1: column tk of the owner's inbox
  // ---- P5 (deferred): the ODE unknowns of the stage just finished.  Every CTA adds the 16 partial moment sets with the same
  // shuffle tree and advances the influencer / media ODEs (pde.jl:239-265) redundantly in Float64: identical results
  // everywhere, no broadcast.  It runs at the start of the NEXT stage's P4 (the only consumer of its results), next to the
  // remote loads of the force halo, instead of on the critical path behind the cluster barrier.
  bool pending = false, lastp = false;
  PdeClusterStage Sp = A.st[0];
  constexpr int TB = NTC - 16 * NV;  // the top 16 * NV threads: (value v, CTA c) = ((tid - TB) / 16, (tid - TB) % 16)
  static_assert(TB >= 0, "too many planes for the moment tree");
  auto ode_sums = [&]() {
    if (tid >= TB) {
      const int q = tid - TB, c = q & 15, v = q >> 4;
      double sum = mom_all[c * NV + v];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (c == 0) sm[v] = sum * dV;  // plane index = j*2+i; entries (mass, x-moment, y-moment)
    }
  };
  auto ode_update = [&]() {
      if (warp == nwarps - 1 && lane < L + 2) {
    const bool is_inf = lane < L;
    const int e = is_inf ? lane : lane - L;
    double* ys = is_inf ? c_y[e] : c_z[e];
    double* dys = is_inf ? c_dy[e] : c_dz[e];
    double* yn = is_inf ? o_yn[e] : o_zn[e];
    double* ya = is_inf ? o_ya[e] : o_za[e];
    double ynew[2];
    for (int dd = 0; dd < 2; dd++) {
      double out = Sp.A * yn[dd] + Sp.Bc * ys[dd] + Sp.C * dt * dys[dd];
      if (Sp.acc == 1) ya[dd] = yn[dd] + dt * Sp.bw * dys[dd];
      else if (Sp.acc == 2) ya[dd] += dt * Sp.bw * dys[dd];
      else if (Sp.acc == 3) out = ya[dd] + dt * Sp.bw * dys[dd];
      ynew[dd] = out;
      if (lastp) yn[dd] = out;
    }
    double m = 0.0, mx = 0.0, my = 0.0;
    if (is_inf) {  // sums over the media i of plane (e, i)
      for (int i = 0; i < 2; i++) { m += sm[(e * 2 + i) * 3]; mx += sm[(e * 2 + i) * 3 + 1]; my += sm[(e * 2 + i) * 3 + 2]; }
    } else {       // sums over the influencers j of plane (j, e)
      for (int j = 0; j < L; j++) { m += sm[(j * 2 + e) * 3]; mx += sm[(j * 2 + e) * 3 + 1]; my += sm[(j * 2 + e) * 3 + 2]; }
    }
    const int ctrl = is_inf ? in.cinf[e] : in.cmed[e];
    double q0, q1;
    if (ctrl == 0 || (!is_inf && ctrl != 1)) {  // free: pde.jl:239-241, :262-264
      const double fr = is_inf ? in.frictionI : in.frictionM;
      q0 = 1.0 / fr * (1.0 / m * mx - ynew[0]);
      q1 = 1.0 / fr * (1.0 / m * my - ynew[1]);
    } else if (ctrl == 1 || (is_inf && ctrl == 3)) {  // controlled: pde.jl:243-261 (tolerance: see finalize_instance, pde.cu)
      const double* tg = ctrl == 1 ? in.tgt1 : in.tgt2;
      const double sp = ctrl == 1 ? in.speed1 : in.speed2;
      const double d0 = tg[0] - ynew[0], d1 = tg[1] - ynew[1];
      const double nrm = sqrt(d0 * d0 + d1 * d1);
      const bool moving = nrm > 1e-12;
      q0 = moving ? sp * d0 / nrm : 0.0;
      q1 = moving ? sp * d1 / nrm : 0.0;
    } else {
      q0 = q1 = 0.0;
    }
    ys[0] = ynew[0]; ys[1] = ynew[1];
    dys[0] = q0; dys[1] = q1;
    if (is_inf) {  // rate factor into influencer e for each medium (pde.jl:43)
      const double m0 = sm[(e * 2 + 0) * 3], m1 = sm[(e * 2 + 1) * 3];
      c_R[0][e] = in.eta * r_fun_c((m0 - m1) / (m0 + m1));
      c_R[1][e] = in.eta * r_fun_c((m1 - m0) / (m1 + m0));
    }
  }
  };
  // ---- P1: row transforms of the density sum (two real rows per complex transform), half spectra to the column owners.
  // It needs this CTA's rows only, so it runs right behind the stencil that produced them, in front of the SAME cluster
  // barrier that publishes the moments: two barriers per stage, not three.
  auto rows_forward = [&]() {
    for (int pr = cP; pr < np; pr += nP) {
      float2 v = make_float2(0.f, 0.f);
      if (tP < Nx) v = make_float2(rho[(2 * pr) * Nx + tP], 2 * pr + 1 < nr ? rho[(2 * pr + 1) * Nx + tP] : 0.f);
      RW[pr * cs + swzP] = v;
    }
    __syncthreads();
    fft_dif_fwd<float>(RW, np, cs, d, tw);
    if (ck < nH)
      for (int pr = ck; pr < np; pr += nH) {
        const float2 zk = RW[pr * cs + tk_pk], zm = RW[pr * cs + tk_pm];
        A_dst[2 * pr] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        if (2 * pr + 1 < nr) A_dst[2 * pr + 1] = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
      }
  };
  int cur = 0;
  rows_forward();
  cluster.sync();

  for (int step = 0; step < A.nsteps; step++) {
    if (pt)  // u_n of this step
      for (int j = 0; j < L; j++) un[(j * RPC + tr) * Nx + tx] = (cur ? U1 : U0)[j * UPS + (tr + 1) * NXP + tx + 1];
    for (int s = 0; s < A.nstages; s++) {
      const PdeClusterStage S = A.st[s];
      const bool last = s == A.nstages - 1;
      float2* Uc = cur ? U1 : U0;
      float2* Un = cur ? U0 : U1;
      DBG(0);
      DBG(3);
      // ---- P2: column transforms, three kernel spectra -------------------------------------------------------------------
      if (pending) ode_sums();  // moments of the previous stage (complete since its cluster barrier)
      for (int c = cP; c < nc; c += nP) R[c * cs + swzP] = tP < Ny ? Arecv[c * Ny + tP] : make_float2(0.f, 0.f);
      __syncthreads();
      fft_dif_fwd<float>(R, nc, cs, d, tw);
      DBG(4);
      for (int k3 = 0; k3 < 3; k3++)
        for (int c = cP; c < nc; c += nP) Wk[(k3 * nc + c) * cs + swzP] = cmul(R[c * cs + swzP], kh[(k3 * CPC + c) * P + tP]);
      __syncthreads();
      fft_dit_inv<float>(Wk, 3 * nc, cs, d, tw);
      DBG(5);
      if (gc < nc && gr < nG)
        for (int iy = gr; iy < Ny; iy += nG) {
          const int owner = iy / RPC, rr = iy - owner * RPC;
          const int src = gc * cs + swz(iy);
          const float2 g0 = Wk[src], g1 = Wk[nc * cs + src], g2 = Wk[2 * nc * cs + src];
          constexpr int GS = (RPC + 2) * hp1;  // stride between the three fields
          float2* dst = cluster.map_shared_rank(Grecv, owner) + (rr + 1) * hp1 + k0 + gc;
          dst[0] = g0; dst[GS] = g1; dst[2 * GS] = g2;
          // the first / last row of a CTA is also the halo row of its neighbour, which transforms it itself (no barrier
          // between the force and the stencil): W rho and the y force only
          if (rr == 0 && owner > 0) {
            float2* h = cluster.map_shared_rank(Grecv, owner - 1) + (RPC + 1) * hp1 + k0 + gc;
            h[0] = g0; h[2 * GS] = g2;
          }
          if (rr == RPC - 1 && iy + 1 < Ny) {
            float2* h = cluster.map_shared_rank(Grecv, owner + 1) + k0 + gc;
            h[0] = g0; h[2 * GS] = g2;
          }
        }
      DBG(6);
      cluster.sync();
      DBG(7);
      // ---- P3: inverse row transforms, force; halo rows of the stage state ------------------------------------------------
      if (pending) ode_update();  // one warp, while the others wait for their remote loads below
      constexpr int GS = (RPC + 2) * hp1;
      const int npe = np + ((has_up || has_dn) ? 1 : 0);  // one more pair: (halo row above, halo row below)
      for (int k3 = 0; k3 < 3; k3++)
        for (int pr = cP; pr < npe; pr += nP) {
          float2 a, bb;
          if (pr < np) {
            a = Grecv[k3 * GS + (2 * pr + 1) * hp1 + fp_src];
            bb = 2 * pr + 1 < nr ? Grecv[k3 * GS + (2 * pr + 2) * hp1 + fp_src] : make_float2(0.f, 0.f);
          } else {
            a = has_up ? Grecv[k3 * GS + fp_src] : make_float2(0.f, 0.f);
            bb = has_dn ? Grecv[k3 * GS + (nr + 1) * hp1 + fp_src] : make_float2(0.f, 0.f);
          }
          if (fp_up) { a.y = -a.y; bb.y = -bb.y; }
          RW[(k3 * npe + pr) * cs + swzP] = make_float2(a.x - bb.y, a.y + bb.x);  // A + i B
        }
      if (tr < nR) {
        const float2* un_ = cur ? upU1 : upU0;
        const float2* dn_ = cur ? dnU1 : dnU0;
        for (int j = tr; j < L; j += nR) {
          if (un_) Uc[j * UPS + tx + 1] = un_[j * UPS + RPC * NXP + tx + 1];
          if (dn_) Uc[j * UPS + (nr + 1) * NXP + tx + 1] = dn_[j * UPS + NXP + tx + 1];
        }
      }
      __syncthreads();
      DBG(8);
      fft_dit_inv<float>(RW, 3 * npe, cs, d, tw);
      DBG(9);
      if (tr < np) {  // thread (tr, tx): row pair tr, column tx
        const float2 o0 = RW[tr * cs + swzX], ox = RW[(npe + tr) * cs + swzX], oy = RW[(2 * npe + tr) * cs + swzX];
        sFx[(2 * tr) * NXP + tx + 1] = ox.x / o0.x;  // pde.jl:58 force ./ norms
        sFy[(2 * tr + 1) * Nx + tx] = oy.x / o0.x;
        if (2 * tr + 1 < nr) {
          sFx[(2 * tr + 1) * NXP + tx + 1] = ox.y / o0.y;
          sFy[(2 * tr + 2) * Nx + tx] = oy.y / o0.y;
        }
      } else if (tr == np && npe > np) {  // the halo pair: y force only
        const float2 o0 = RW[np * cs + swzX], oy = RW[(2 * npe + np) * cs + swzX];
        if (has_up) sFy[tx] = oy.x / o0.x;
        if (has_dn) sFy[(nr + 1) * Nx + tx] = oy.y / o0.y;
      }
      DBG(10);
      __syncthreads();
      DBG(11);
      // ---- P4: right-hand side and RK update of this CTA's rows ----------------------------------------------------------
      if (pending) { refresh_f32(); __syncthreads(); }
      const float dtA = (float)S.A, dtB = (float)S.Bc, dtC = (float)(S.C * dt), dtw = (float)(S.bw * dt);
      if (pt) {
        const int r = tr, x = tx;
        const float* fx = sFx + r * NXP + x + 1;
        const float* fy = sFy + (r + 1) * Nx + x;
        // plane-independent parts of the velocity a F + b (y_j - x) + c (z_i - x)   (pde.jl:214)
        const float vxc0 = pa * fx[0] - bc * X, vxm0 = pa * fx[-1] - bc * (X - h), vxp0 = pa * fx[1] - bc * (X + h);
        const float vyc0 = pa * fy[0] - bc * Y, vym0 = pa * fy[-Nx] - bc * (Y - h), vyp0 = pa * fy[Nx] - bc * (Y + h);
        const float2* tc = Uc + (r + 1) * NXP + x + 1;  // centre, influencer 0
        const float2 m1 = make_float2(-1.f, -1.f);
        auto sub2 = [&](float2 a_, float2 b_) { return fma2(b_, m1, a_); };
        auto bc2 = [](float v) { return make_float2(v, v); };
        float2 rsum = make_float2(0.f, 0.f), gsum = rsum;   // per medium: sum of densities and of rates (pde.jl:220-225)
        float E[LT];
#pragma unroll
        for (int j = 0; j < L; j++) {
          const float ex = X - s_yx[j], ey = Y - s_yy[j];
          E[j] = w_of_r2(fmaf(ex, ex, ey * ey));
          rsum = add2(rsum, tc[j * UPS]);
          gsum = fma2(bc2(E[j]), s_R2[j], gsum);
        }
        float2 rs2 = make_float2(0.f, 0.f);
        const float2 hih = bc2(0.5f * ih), ih_2 = bc2(ih), ih2_2 = bc2(ih2), Dd2 = bc2(Ddx), D2 = bc2(D);
#pragma unroll
        for (int j = 0; j < L; j++) {
          const float2* tp = tc + j * UPS;
          const float2 c = tp[0], vxm = tp[-1], vxp = tp[1], vym = tp[-NXP], vyp = tp[NXP];
          const float2 bxv = s_bx2[j], byv = s_by2[j];  // b y_j + c z_i for the two media
          const float2 fxc = mul2(c, add2(bc2(vxc0), bxv)), fxm = mul2(vxm, add2(bc2(vxm0), bxv)), fxp = mul2(vxp, add2(bc2(vxp0), bxv));
          const float2 fyc = mul2(c, add2(bc2(vyc0), byv)), fym = mul2(vym, add2(bc2(vym0), byv)), fyp = mul2(vyp, add2(bc2(vyp0), byv));
          // divergence with C (pde.jl:79-87,216-217), diffusion with MassM (pde.jl:68-77,228)
          float2 ddx = mul2(sub2(fxp, fxm), hih), lapx = mul2(fma2(bc2(-2.f), c, add2(vxm, vxp)), ih2_2);
          float2 ddy = mul2(sub2(fyp, fym), hih), lapy = mul2(fma2(bc2(-2.f), c, add2(vym, vyp)), ih2_2), bnd = make_float2(0.f, 0.f);
          if (x_lo) { ddx = mul2(sub2(fxp, fxc), ih_2); lapx = mul2(sub2(vxp, c), ih2_2); bnd = sub2(bnd, mul2(Dd2, fxc)); }
          if (x_hi) { ddx = mul2(sub2(fxc, fxm), ih_2); lapx = mul2(sub2(vxm, c), ih2_2); bnd = fma2(Dd2, fxc, bnd); }
          if (y_lo) { ddy = mul2(sub2(fyp, fyc), ih_2); lapy = mul2(sub2(vyp, c), ih2_2); bnd = sub2(bnd, mul2(Dd2, fyc)); }
          if (y_hi) { ddy = mul2(sub2(fyc, fym), ih_2); lapy = mul2(sub2(vym, c), ih2_2); bnd = fma2(Dd2, fyc, bnd); }
          const float2 G = mul2(bc2(E[j]), s_R2[j]);
          const float2 reac = sub2(mul2(G, sub2(rsum, c)), mul2(c, sub2(gsum, G)));
          const float2 k = add2(sub2(fma2(D2, add2(lapx, lapy), bnd), add2(ddx, ddy)), reac);  // pde.jl:236
          const int o = (j * RPC + r) * Nx + x;
          const float2 unp = un[o];
          float2 out = mul2(bc2(dtC), k);
          if (dtA != 0.f) out = fma2(bc2(dtA), unp, out);
          if (dtB != 0.f) out = fma2(bc2(dtB), c, out);
          if (S.acc == 1) acc[o] = fma2(bc2(dtw), k, unp);            // ACC_INIT
          else if (S.acc == 2) acc[o] = fma2(bc2(dtw), k, acc[o]);    // ACC_ADD
          else if (S.acc == 3) out = fma2(bc2(dtw), k, acc[o]);       // ACC_FINAL
          Un[j * UPS + (r + 1) * NXP + x + 1] = out;
          rs2 = add2(rs2, out);
        }
        const float rs_out = rs2.x + rs2.y;
        rho[r * Nx + x] = rs_out;
      }
      __syncthreads();
      DBG(12);
      // moments of the new stage state: mass, x-moment, y-moment per plane over this CTA's rows
      for (int p = warp; p < NP; p += nwarps) {
        float s0 = 0.f, sx = 0.f, sy = 0.f;
        for (int x = lane; x < Nx; x += 32) {
          float cs0 = 0.f, csy = 0.f;
          for (int r = 0; r < nr; r++) {
            const float2 v2 = Un[(p >> 1) * UPS + (r + 1) * NXP + x + 1];
            const float v = (p & 1) ? v2.y : v2.x;
            cs0 += v;
            csy += v * (float)(A.gy0 + (r0 + r) * A.dx);
          }
          s0 += cs0; sy += csy; sx += cs0 * (float)(A.gx0 + x * A.dx);
        }
        s0 = warp_sum(s0); sx = warp_sum(sx); sy = warp_sum(sy);
        if (lane == 0) { part[p * 3] = (double)s0; part[p * 3 + 1] = (double)sx; part[p * 3 + 2] = (double)sy; }
      }
      __syncthreads();
      for (int e = tid; e < CS * NV; e += NTC) {
        const int dest = e / NV, v = e - dest * NV;
        cluster.map_shared_rank(mom_all, dest)[rank * NV + v] = part[v];
      }
      DBG(13);
      if (!(last && step == A.nsteps - 1)) rows_forward();  // P1 of the next stage
      cluster.sync();
      DBG(14);
      pending = true; Sp = S; lastp = last;
      DBG(15);
      cur ^= 1;
    }
  }
  if (pending) {
    ode_sums();
    __syncthreads();
    ode_update();
    __syncthreads();
  }
  // ---- store -------------------------------------------------------------------------------------------------------------
  if (pt) {
    for (int j = 0; j < L; j++) {
      const float2 v = (cur ? U1 : U0)[j * UPS + (tr + 1) * NXP + tx + 1];
      const size_t o = (size_t)(r0 + tr) * A.Pw + tx;
      ug[(size_t)(2 * j) * plane + o] = v.x; ug[(size_t)(2 * j + 1) * plane + o] = v.y;
    }
    rhog[(size_t)(r0 + tr) * A.Pw + tx] = rho[tr * Nx + tx];
  }
  if (rank == 0) {
    if (tid < L) {
      for (int dd = 0; dd < 2; dd++) {
        cfg.y[tid][dd] = c_y[tid][dd]; cfg.dy[tid][dd] = c_dy[tid][dd]; cfg.R[dd][tid] = c_R[dd][tid];
        odg.yn[tid][dd] = o_yn[tid][dd]; odg.yacc[tid][dd] = o_ya[tid][dd];
      }
    }
    if (tid < 2) {
      for (int dd = 0; dd < 2; dd++) {
        cfg.z[tid][dd] = c_z[tid][dd]; cfg.dz[tid][dd] = c_dz[tid][dd];
        odg.zn[tid][dd] = o_zn[tid][dd]; odg.zacc[tid][dd] = o_za[tid][dd];
      }
    }
  }
  cluster.sync();  // no CTA leaves while a neighbour may still read its shared memory
}

using ClusterKernel = void (*)(const PdeClusterArgs);

// The shapes the reference constructs (setting.jl:45-75): PDEconstruct 89 x 89 (P = 192), PDEconstructcoarse 47 x 47 (P = 96);
// L = 4 (+ 1 or 2 controlled influencers, controlstrategies.jl).  Rows / columns per CTA: ceil over the 16 CTAs, rows even.
ClusterKernel kernel_for(int N, int L, int P, int* RPC, int* CPC) {
  if (N == 89 && P == 192) {
    *RPC = 6; *CPC = 7;
    if (L == 4) return pde_cluster_kernel<4, 89, 192, 6, 7>;
    if (L == 5) return pde_cluster_kernel<5, 89, 192, 6, 7>;
  }
  if (N == 47 && P == 96) {
    *RPC = 4; *CPC = 4;
    if (L == 4) return pde_cluster_kernel<4, 47, 96, 4, 4>;
    if (L == 5) return pde_cluster_kernel<5, 47, 96, 4, 4>;
    if (L == 6) return pde_cluster_kernel<6, 47, 96, 4, 4>;
  }
  return nullptr;
}

void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int B, size_t smem, cudaStream_t st) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(CS, B); cfg->blockDim = dim3(NTC); cfg->dynamicSmemBytes = smem; cfg->stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg->attrs = at; cfg->numAttrs = 1;
}

}  // namespace

// Can the cluster-resident form run this instance shape?  (FP32, a shape instantiated above, shared memory, and a 16-CTA
// cluster schedulable on this device)
int smm_pde_cluster_plan(smm_ctx* ctx, int Nx, int Ny, int L, int P, int* RPC, int* CPC, size_t* smem) {
  *smem = 0;
  if (Nx != Ny) return SMM_ERR_UNSUPPORTED;
  ClusterKernel k = kernel_for(Nx, L, P, RPC, CPC);
  if (!k) return SMM_ERR_UNSUPPORTED;
  const Lay l = make_lay(L, Nx, Ny, P, *RPC, *CPC);
  if (l.end + 2048 > 227 * 1024) return SMM_ERR_UNSUPPORTED;
  int clusters = 0;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  cluster_config(&cfg, at, 1, l.end, nullptr);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.end) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&clusters, k, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return SMM_ERR_UNSUPPORTED;
  }
  *smem = l.end;
  return SMM_OK;
}

int smm_launch_pde_cluster(smm_ctx* ctx, const PdeClusterArgs& A, int B, size_t smem) {
  int rpc, cpc;
  ClusterKernel k = kernel_for(A.Nx, A.L, A.P, &rpc, &cpc);
  SMM_REQUIRE(ctx, k && rpc == A.RPC && cpc == A.CPC, "smm_launch_pde_cluster: shape %d x %d, L = %d not instantiated", A.Nx, A.Ny, A.L);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  cluster_config(&cfg, at, B, smem, ctx->stream);
  SMM_CUDA(ctx, cudaLaunchKernelEx(&cfg, k, A));
  ctx->launches++;
  return SMM_OK;
}

#ifdef SMM_CLUSTER_DEBUG
extern "C" int smm_debug_cluster_times(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 64);
}
#endif
