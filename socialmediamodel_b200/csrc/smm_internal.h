// smm_internal.h -- argument blocks and launchers shared between the .cu files of libsmm_b200.so
#pragma once
#include "smm_common.cuh"

// K1 dense ensemble kernel (abm_dense.cu). All pointers are DEVICE pointers; layouts are the
// Julia layouts documented in include/smm_b200.h with `real` = float|double for x0/xi/xs.
struct AbmDenseArgs {
  AbmConst k;
  int NT, R;
  int64_t first_realization;
  const void* x0;           // real  [R][2][n]
  const signed char* state; // int8  [R][n]
  const int32_t* follow0;   // int32 [R][n]
  const double* inf0;       // f64   [R][2][L]
  const double* media0;     // f64   [R][2][M]
  const int32_t* rowptr;    // CSR network shared by all realizations, or null = all-to-all
  const int32_t* colidx;
  int noise_mode;
  uint64_t seed;
  const void* xi;           // real  [R][NT][2][n]
  const double* u1;         // f64   [R][NT][n]
  const double* u2;
  const double* xi_inf;     // f64   [R][NT][L][2] or null
  const double* xi_med;     // f64   [R][NT][M][2] or null
  int nsave;
  const int32_t* savesteps; // device int32 [nsave]
  void* xs;                 // real  [R][nsave][2][n]
  int32_t* follows;         // int32 [R][nsave][n]
  double* infs;             // f64   [R][NT+1][2][L]
  double* meds;             // f64   [R][NT+1][2][M]
  double* ys;               // f64   [R][nsave][L][2]
  double* zs;               // f64   [R][nsave][M][2]
  int64_t* numbers;         // int64 [NT+1][L][M] (atomically accumulated)
  // KDE of the save steps inside the kernel (FP32 symmetric path; api_abm.cu asks smm_abm_dense_fuses_kde first):
  // kde_us = real [R][nsave][L][2][Ny][Nx] or null (then xs / follows feed kde_kernel)
  void* kde_us;
  int kde_Nx, kde_Ny;
  double kde_gx0, kde_gy0, kde_dx, kde_dy, kde_sigma;
  // set by smm_launch_abm_dense (step chunks, see abm_dense.cu): state handed from chunk to chunk, chunks done per realization
  int chunk_steps;          // 0 = the whole run in one CTA
  int p2_cost;              // wide CTAs: head start of warp 0 in the tile dealing, in tile steps (it also updates the influencers)
  void* hand_x;             // real  [R][2][n]
  int32_t* hand_f;          // int32 [R][n]
  double* hand_inf;         // f64   [R][2][L]
  double* hand_med;         // f64   [R][2][M]
  unsigned* chunk_done;     // [R]
};

struct AbmInitArgs {
  AbmConst k;
  int cells;  // L = cells*cells
  uint64_t seed;
  int64_t first_realization;
  void* x0;
  signed char* state;
  int32_t* follow0;
  double* inf0;
  double* media0;
};

int smm_launch_abm_dense(smm_ctx* ctx, const AbmDenseArgs& A, int precision);
int smm_launch_abm_init(smm_ctx* ctx, const AbmInitArgs& A, int R, int precision);
int smm_launch_abm_draws(smm_ctx* ctx, int n, uint64_t seed, uint32_t realization, int NT, double* xi, double* u1, double* u2);

// K4 KDE of snapshots (kde.cu): us[s][j][i][iy][ix] for S = R*nsave snapshots
struct KdeArgs {
  int n, L, S;
  int Nx, Ny;
  double gx0, gy0, dx, dy, sigma;
  const void* xs;               // real [S][2][n]
  const int32_t* follows;       // [S][n]
  const signed char* state;     // [R][n], snapshot s belongs to realization s / nsave
  int nsave;
  void* us;                     // real [S][L][2][Ny][Nx]
};
int smm_launch_kde(smm_ctx* ctx, const KdeArgs& A, int precision);
// true when smm_launch_abm_dense can evaluate the KDE of the save steps itself for this configuration
bool smm_abm_dense_fuses_kde(const smm_ctx* ctx, const AbmDenseArgs& A, int precision, int Nx, int Ny);
// K4b ensemble reduction: sum over realizations (plotting.jl:361-363 numerator), f64 output
int smm_launch_ensemble_sum(smm_ctx* ctx, const void* us, int precision, int R, size_t per_real, double* sum_us, int accumulate = 0);
int smm_launch_small_sum(smm_ctx* ctx, const double* v, int R, int per_real, double* out);
// element conversions
int smm_launch_convert_f64_to_real(smm_ctx* ctx, const double* in, void* out, size_t count, int precision);
int smm_launch_convert_real_to_f64(smm_ctx* ctx, const void* in, double* out, size_t count, int precision);
int smm_launch_convert_i32_to_i8(smm_ctx* ctx, const int32_t* in, signed char* out, size_t count);

// CSR population kernels (abm_csr.cu)
struct smm_abm_pop {
  AbmConst k;
  int precision;
  int device;        // device whose persisting-L2 carve-out this population holds a share of (-1: none)
  uint64_t seed;
  int64_t step;      // number of steps taken so far
  int has_net;       // 0 = all-to-all
  size_t nnz;
  double fx_scale;   // fixed-point scale of the position accumulators (power of two)
  void* x[2];        // real2 [n], double-buffered (step parity)
  unsigned char* follow;  // [n]
  signed char* state;     // [n]
  int32_t* rowptr;
  int32_t* colidx;     // SELL-32 layout (see abm_csr.cu)
  int64_t* slice_ptr;  // [n/32 + 1] offsets into colidx
  double* ent[2];    // [2][L+M] influencer then media positions (x block, y block), double-buffered
  long long* acc[3]; // fixed-point class sums [2L][3] = (sum x, sum y, count), 3-way rotation
  unsigned long long* fallback_count;  // FP64 re-evaluations of the switch decision so far
  int prefix_decision;  // switch-decision variant (SMM_POP_DECISION=prefix|coop)
  size_t l2_persist_bytes, l2_window_max;  // L2 persistence for the gathered position array (0 = off)
};
int smm_launch_pop_accumulate(smm_ctx* ctx, smm_abm_pop* p, int buf, int accslot);
int smm_launch_pop_step(smm_ctx* ctx, smm_abm_pop* p, const void* xi, const double* u1, const double* u2);
int smm_launch_csr_to_sell(smm_ctx* ctx, int n, const int32_t* rowptr, const int32_t* csr, const int64_t* slice_ptr, int32_t* sell);
int smm_launch_pack_xy(smm_ctx* ctx, const double* xy, void* x2, size_t n, int precision);    // [2][n] f64 -> real2[n]
int smm_launch_unpack_xy(smm_ctx* ctx, const void* x2, double* xy, size_t n, int precision);  // real2[n] -> [2][n] f64
int smm_launch_convert_i32_to_u8(smm_ctx* ctx, const int32_t* in, unsigned char* out, size_t count);
int smm_launch_convert_u8_to_i32(smm_ctx* ctx, const unsigned char* in, int32_t* out, size_t count);

// FFT-based Toeplitz correlation for the PDE interaction force (fft_conv.cu)
struct smm_fft_plan;
int smm_fft_plan_create(smm_ctx* ctx, int Nx, int Ny, int Pw, double hx, double hy, int B, int precision, smm_fft_plan** out);
int smm_fft_conv(smm_ctx* ctx, smm_fft_plan* p, const void* rhosum, void* Fx, void* Fy);
void smm_fft_plan_destroy(smm_fft_plan* p);
void smm_fft_plan_tables(const smm_fft_plan* p, int dims[4] /* P, m, a, r3 */, const void** tw, const void** khat);
// whole RK steps of small instances inside one thread-block cluster (pde_cluster.cu)
struct PdeClusterStage { double A, Bc, C, bw; int acc; };
struct PdeClusterArgs {
  int L, Nx, Ny, Pw, P, m, a, r3;  // grid, row pitch, FFT length P = r3 * m, m = 2^a
  int RPC, CPC, nsteps, nstages;   // grid rows / spectrum columns per CTA
  double gx0, gy0, dx, dt;
  float* u;                        // [B][2L][Ny][Pw]
  float* rhosum;                   // [B][Ny][Pw]
  void* coef;                      // PdeCoef[B]
  void* ode;                       // PdeOde[B]
  const void* inst;                // PdeInst[B]
  const void* tw;                  // float2[P]
  const void* khat;                // float2[3][P/2+1][P]
  PdeClusterStage st[4];
};
int smm_pde_cluster_plan(smm_ctx* ctx, int Nx, int Ny, int L, int P, int* RPC, int* CPC, size_t* smem);
int smm_launch_pde_cluster(smm_ctx* ctx, const PdeClusterArgs& A, int B, size_t smem);
// register-resident FP32 path for P = 96 / 192 / 384 / 768 (fft_fast.cu)
struct smm_fft_fast;
bool smm_fft_fast_supported(int P);
int smm_fft_fast_create(smm_ctx* ctx, int P, int Nx, int Ny, int Pw, int B, const void* khat64, const int* slot_of_freq_dev, smm_fft_fast** out);
int smm_fft_fast_conv(smm_ctx* ctx, smm_fft_fast* f, const void* rhosum, void* Fx, void* Fy);
void smm_fft_fast_destroy(smm_fft_fast* f);
double smm_fft_fast_residual(const smm_fft_fast* f);

// several GPUs behind one context (multi.cu)
struct smm_multi;
void smm_multi_destroy(smm_multi* m);
int smm_abm_ensemble_shard(smm_ctx* ctx, const smm_params* q, int32_t precision, int32_t NT, int32_t R, int64_t first_realization,
                           const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0, const double* media0,
                           const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                           const int32_t* savesteps, double kde_sigma, const smm_grid* grid, void* us, bool us_f32, double* zs,
                           double* ys, double* sum_us, double* sum_zs, double* sum_ys, int64_t* numbers);
int smm_multi_abm_ensemble(smm_ctx* lead, const smm_params* q, int32_t precision, int32_t NT, int32_t R, int64_t first,
                           const double* x0, const int32_t* state, const int32_t* follow0, const double* inf0, const double* media0,
                           const int32_t* rowptr, const int32_t* colidx, const smm_noise* noise, int32_t nsave,
                           const int32_t* savesteps, double kde_sigma, const smm_grid* grid, void* us, bool us_f32, double* zs,
                           double* ys, double* sum_us, double* sum_zs, double* sum_ys, int64_t* numbers);
int smm_multi_pde_solve(smm_ctx* lead, const smm_grid* grid, const smm_params* q, int32_t B, int32_t L, int32_t precision,
                        const double* controlled_inf, const double* controlled_med, const double* u0, const double* z0,
                        const double* y0, double dt, int32_t nsteps, int32_t scheme, int32_t nsave, const int32_t* savesteps,
                        double* us, double* zs, double* ys, double* sum_us, double* sum_zs, double* sum_ys);
