/* abm_oracle.c -- plain C restatement of the reference's ABM ensemble path, Float64 throughout.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): the checker and the CPU baseline of bench.py, never
 * the product.  PARITY UNPINNED: the reference ships no golden vectors and Julia is not available;
 * this file is validated against the NumPy restatement (tests/test_oracle_c.py), which is pinned to
 * the closed-form known answers of tests/test_oracle_kats.py.
 *
 * Restates, loop for loop (so that its cost is the cost of the reference's algorithm on a CPU):
 *   attraction        src/abm.jl:47-67     all-to-all incl. self, sequential sums over i
 *   influence         src/abm.jl:69-88
 *   changeinfluencer  src/abm.jl:90-137    with the updated influencers (aliasing, :175,207)
 *   boundary          src/abm.jl:139-150   single reflection
 *   loop body         src/abm.jl:166-213
 *   KDE               src/ensemble.jl:2-8,27-41 and gaussian src/pde.jl:122: for every grid point the
 *                     2-D Gaussian of every agent of the class (NOT separated into 1-D factors)
 * Arrays are C row-major: x[n][2], inf[L][2], media[M][2]; follow in 0..L-1; state in {-1,+1}.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n, L, M;
  double a, b, c, eta, sigma, sigmahat, sigmatilde, frictionI, frictionM, dt;
  double lo[2], hi[2];
} smm_c_params;

static void reflect(double* v, int count, const smm_c_params* q) { /* abm.jl:139-150 */
  for (int j = 0; j < count; j++)
    for (int d = 0; d < 2; d++) {
      double* p = &v[2 * j + d];
      if (*p > q->hi[d]) *p = -*p + 2 * q->hi[d];
      else if (*p < q->lo[d]) *p = -*p + 2 * q->lo[d];
    }
}

/* one pass of abm.jl:166-213; xi[n][2], u1[n], u2[n]; work: 2n doubles; follow/x/inf/media updated in place */
static void step(const smm_c_params* q, double* x, double* inf, double* media, int* follow, const int* state,
                 const double* xi, const double* u1, const double* u2, const double* xi_inf, const double* xi_med,
                 double* xnew) {
  const int n = q->n, L = q->L, M = q->M;
  const double dt = q->dt, sdt = sqrt(dt);
  for (int j = 0; j < n; j++) { /* abm.jl:47-88, :173 */
    double fx = 0, fy = 0, ws = 0;
    for (int i = 0; i < n; i++) {
      const double dx = x[2 * i] - x[2 * j], dy = x[2 * i + 1] - x[2 * j + 1];
      const double w = exp(-sqrt(dx * dx + dy * dy));
      fx += w * dx; fy += w * dy; ws += w;
    }
    const int m = state[j] == 1 ? 1 : 0, f = follow[j];
    const double Fx = q->a * (fx / ws) + q->c * (media[2 * m] - x[2 * j]) + q->b * (inf[2 * f] - x[2 * j]);
    const double Fy = q->a * (fy / ws) + q->c * (media[2 * m + 1] - x[2 * j + 1]) + q->b * (inf[2 * f + 1] - x[2 * j + 1]);
    xnew[2 * j] = x[2 * j] + dt * Fx + sdt * q->sigma * xi[2 * j];
    xnew[2 * j + 1] = x[2 * j + 1] + dt * Fy + sdt * q->sigma * xi[2 * j + 1];
  }
  /* influencers and media from the OLD positions / followers (abm.jl:175-198) */
  double frac[256];
  for (int i = 0; i < L; i++) {
    double sx = 0, sy = 0, ss = 0;
    int cnt = 0;
    for (int j = 0; j < n; j++)
      if (follow[j] == i) { sx += x[2 * j]; sy += x[2 * j + 1]; ss += state[j]; cnt++; }
    const double nz = 1.0 / q->frictionI * sdt * q->sigmahat;
    const double z0 = xi_inf ? xi_inf[2 * i] : 0.0, z1 = xi_inf ? xi_inf[2 * i + 1] : 0.0;
    if (cnt > 0) {
      inf[2 * i] = inf[2 * i] + dt / q->frictionI * (sx / cnt - inf[2 * i]) + nz * z0;
      inf[2 * i + 1] = inf[2 * i + 1] + dt / q->frictionI * (sy / cnt - inf[2 * i + 1]) + nz * z1;
    } else {
      inf[2 * i] += nz * z0; inf[2 * i + 1] += nz * z1;
    }
    frac[i] = ss / (double)cnt; /* abm.jl:99: 0/0 = NaN */
  }
  for (int i = 0; i < M; i++) {
    const int s = i == 0 ? -1 : 1;
    double sx = 0, sy = 0;
    int cnt = 0;
    for (int j = 0; j < n; j++)
      if (state[j] == s) { sx += x[2 * j]; sy += x[2 * j + 1]; cnt++; }
    const double nz = 1.0 / q->frictionM * sdt * q->sigmatilde;
    const double z0 = xi_med ? xi_med[2 * i] : 0.0, z1 = xi_med ? xi_med[2 * i + 1] : 0.0;
    if (cnt > 0) {
      media[2 * i] = media[2 * i] + dt / q->frictionM * (sx / cnt - media[2 * i]) + nz * z0;
      media[2 * i + 1] = media[2 * i + 1] + dt / q->frictionM * (sy / cnt - media[2 * i + 1]) + nz * z1;
    } else {
      media[2 * i] += nz * z0; media[2 * i + 1] += nz * z1;
    }
  }
  reflect(xnew, n, q); reflect(inf, L, q); reflect(media, M, q); /* abm.jl:201-203 */
  /* changeinfluencer with OLD x, OLD followers' fractions, UPDATED influencers (abm.jl:90-137, :207) */
  double att[256];
  for (int j = 0; j < n; j++) {
    double lam = 0;
    for (int i = 0; i < L; i++) {
      const double dx = x[2 * j] - inf[2 * i], dy = x[2 * j + 1] - inf[2 * i + 1];
      double g2 = state[j] * frac[i];
      if (g2 < 0.1) g2 = 0.1; /* NaN < 0.1 is false */
      att[i] = q->eta * exp(-sqrt(dx * dx + dy * dy)) * g2;
      lam += att[i];
    }
    if (u1[j] < 1.0 - exp(-lam * dt)) { /* NaN -> false */
      double cs = 0;
      int k = L - 1;
      for (int i = 0; i < L; i++) {
        cs += att[i] / lam;
        if (!(cs < u2[j])) { k = i; break; }
      }
      follow[j] = k;
    }
  }
  memcpy(x, xnew, sizeof(double) * 2 * n);
}

/* ABMsolve with injected initial condition and draws (abm.jl:154-217): histories xs[NT+1][n][2],
 * follows[NT+1][n], infs[NT+1][L][2], meds[NT+1][M][2]; xi[NT][n][2], u1[NT][n], u2[NT][n]. */
int smm_c_abm_solve(const smm_c_params* q, int NT, const double* x0, const int* state, const int* follow0,
                    const double* inf0, const double* media0, const double* xi, const double* u1, const double* u2,
                    double* xs, int* follows, double* infs, double* meds) {
  const int n = q->n, L = q->L, M = q->M;
  if (L > 256) return -1;
  double* x = malloc(sizeof(double) * 2 * n), *xn = malloc(sizeof(double) * 2 * n);
  double* inf = malloc(sizeof(double) * 2 * L), *media = malloc(sizeof(double) * 2 * M);
  int* follow = malloc(sizeof(int) * n);
  memcpy(x, x0, sizeof(double) * 2 * n); memcpy(inf, inf0, sizeof(double) * 2 * L);
  memcpy(media, media0, sizeof(double) * 2 * M); memcpy(follow, follow0, sizeof(int) * n);
  for (int k = 0; k <= NT; k++) {
    memcpy(xs + (size_t)k * 2 * n, x, sizeof(double) * 2 * n);
    memcpy(follows + (size_t)k * n, follow, sizeof(int) * n);
    memcpy(infs + (size_t)k * 2 * L, inf, sizeof(double) * 2 * L);
    memcpy(meds + (size_t)k * 2 * M, media, sizeof(double) * 2 * M);
    if (k == NT) break;
    step(q, x, inf, media, follow, state, xi + (size_t)k * 2 * n, u1 + (size_t)k * n, u2 + (size_t)k * n, NULL, NULL, xn);
  }
  free(x); free(xn); free(inf); free(media); free(follow);
  return 0;
}

/* ensemble.jl:27-38: us[i][j][ix][iy] (medium i, influencer j) = (1/n) sum over the agents of the class of
 * gaussian([X,Y], x_agent) with gaussian(x, c) = 1/(2 pi sigma^2) exp(-|x-c|^2 / (2 sigma^2)) (pde.jl:122). */
void smm_c_kde(const smm_c_params* q, const double* x, const int* follow, const int* state, int Nx, int Ny, double gx0,
               double gy0, double dx, double dy, double sigma, double* us) {
  const int n = q->n, L = q->L;
  const double pre = 1.0 / (2.0 * M_PI * sigma * sigma), inv = 1.0 / (2.0 * sigma * sigma);
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < L; j++) {
      double* o = us + ((size_t)i * L + j) * Nx * Ny;
      for (int ix = 0; ix < Nx; ix++)
        for (int iy = 0; iy < Ny; iy++) {
          const double X = gx0 + ix * dx, Y = gy0 + iy * dy;
          double s = 0;
          for (int a = 0; a < n; a++)
            if (follow[a] == j && state[a] == (i == 0 ? -1 : 1)) {
              const double ex = X - x[2 * a], ey = Y - x[2 * a + 1];
              s += pre * exp(-inv * (ex * ex + ey * ey));
            }
          o[(size_t)ix * Ny + iy] = (1.0 / n) * s;
        }
    }
}

/* ---- throughput run for bench.py: ABMsolveensemble (ensemble.jl:11-55) on R realizations ------------ */
static inline uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double urand(uint64_t* s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline void nrand2(uint64_t* s, double* a, double* b) { /* Box-Muller */
  const double u = 1.0 - urand(s), v = urand(s), r = sqrt(-2.0 * log(u));
  *a = r * cos(2.0 * M_PI * v); *b = r * sin(2.0 * M_PI * v);
}

/* Runs realizations first..first+R-1 (seed = index, ensemble.jl:25) with NT steps and a KDE at the nsave save
 * steps; returns the sum of all density values (a checksum that keeps the work alive).  threads <= 0: all cores. */
double smm_c_ensemble_bench(const smm_c_params* q, int NT, int R, int64_t first, int nsave, const int* savesteps, int Nx,
                            int Ny, double gx0, double gy0, double dx, double dy, double sigma, int threads, int* threads_used) {
  const int n = q->n, L = q->L, M = q->M;
  double total = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  if (threads_used) *threads_used = omp_get_max_threads();
#else
  if (threads_used) *threads_used = 1;
#endif
#pragma omp parallel reduction(+ : total)
  {
    double* x = malloc(sizeof(double) * 2 * n), *xn = malloc(sizeof(double) * 2 * n), *xi = malloc(sizeof(double) * 2 * n);
    double* u1 = malloc(sizeof(double) * n), *u2 = malloc(sizeof(double) * n);
    double* inf = malloc(sizeof(double) * 2 * L), *media = malloc(sizeof(double) * 2 * M);
    int* follow = malloc(sizeof(int) * n), *state = malloc(sizeof(int) * n);
    double* us = malloc(sizeof(double) * 2 * L * Nx * Ny);
#pragma omp for schedule(dynamic, 1)
    for (int r = 0; r < R; r++) {
      uint64_t s = 0x1234567ull + (uint64_t)(first + r) * 0x9E3779B97F4A7C15ull;
      /* ABMinit (abm.jl:8-45): uniform positions, quadrant followers, influencer = follower mean, media +-1 */
      int cnt[4] = {0, 0, 0, 0};
      for (int i = 0; i < 2 * L; i++) inf[i] = 0;
      for (int j = 0; j < n; j++) {
        x[2 * j] = q->lo[0] + (q->hi[0] - q->lo[0]) * urand(&s);
        x[2 * j + 1] = q->lo[1] + (q->hi[1] - q->lo[1]) * urand(&s);
        state[j] = urand(&s) > 0.5 ? 1 : -1;
        const int f = x[2 * j + 1] > 0 ? (x[2 * j] > 0 ? 0 : 1) : (x[2 * j] > 0 ? 2 : 3);
        follow[j] = f % L;
        inf[2 * follow[j]] += x[2 * j]; inf[2 * follow[j] + 1] += x[2 * j + 1]; cnt[follow[j] % 4]++;
      }
      for (int i = 0; i < L && i < 4; i++) { inf[2 * i] /= cnt[i]; inf[2 * i + 1] /= cnt[i]; }
      for (int i = 0; i < M; i++) { media[2 * i] = i == 0 ? -1.0 : 1.0; media[2 * i + 1] = media[2 * i]; }
      int next = 0;
      for (int k = 0; k <= NT; k++) {
        if (next < nsave && savesteps[next] == k) {
          smm_c_kde(q, x, follow, state, Nx, Ny, gx0, gy0, dx, dy, sigma, us);
          for (size_t e = 0; e < (size_t)2 * L * Nx * Ny; e++) total += us[e];
          next++;
        }
        if (k == NT) break;
        for (int j = 0; j < n; j++) { nrand2(&s, &xi[2 * j], &xi[2 * j + 1]); u1[j] = urand(&s); u2[j] = urand(&s); }
        step(q, x, inf, media, follow, state, xi, u1, u2, NULL, NULL, xn);
      }
    }
    free(x); free(xn); free(xi); free(u1); free(u2); free(inf); free(media); free(follow); free(state); free(us);
  }
  return total;
}
