/* A plain C99 caller of the C ABI (include/smm_b200.h): the binding a compiled host language makes.
 * Runs a small ABMsolveensemble and prints a checksum the Python test compares with the ctypes path.
 * Build: gcc -std=c99 -I include tests/c/abi_example.c -o abi_example -L socialmediamodel_b200 -lsmm_b200 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smm_b200.h"

int main(void) {
  smm_ctx* ctx = NULL;
  if (smm_ctx_create(0, &ctx) != SMM_OK) { fprintf(stderr, "ctx: %s\n", smm_last_error(NULL)); return 2; }
  smm_params q;
  memset(&q, 0, sizeof q);
  q.n = 50; q.L = 4; q.M = 2;                       /* setting.jl:1-22 defaults, fewer agents */
  q.eta = 15.0; q.a = 1.0; q.b = 4.0; q.c = 2.0; q.sigma = 0.5; q.frictionI = 10.0; q.frictionM = 100.0;
  q.dt = 0.01; q.dom_lo[0] = q.dom_lo[1] = -2.2; q.dom_hi[0] = q.dom_hi[1] = 2.2;
  smm_noise ns;
  memset(&ns, 0, sizeof ns);
  ns.mode = SMM_NOISE_PHILOX; ns.seed = 0;
  smm_grid g = {23, 23, -2.2, -2.2, 0.2, 0.2};      /* ABMconstruct(dx = 0.2) */
  const int NT = 20, R = 4, nsave = 2;
  int32_t savesteps[2] = {0, 19};                   /* Int.(round.(LinRange(1, NT, 2))) - 1 */
  size_t per = (size_t)g.Nx * g.Ny * 2 * q.L * nsave;
  double* us = (double*)malloc(per * R * sizeof(double));
  double sum_ys[2 * 4 * 2];
  int rc = smm_abm_ensemble(ctx, &q, SMM_PREC_F32, NT, R, 1, NULL, NULL, NULL, NULL, NULL, NULL, NULL, &ns, nsave, savesteps,
                            0.1, &g, us, NULL, NULL, NULL, NULL, sum_ys, NULL);
  if (rc != SMM_OK) { fprintf(stderr, "ensemble: %s\n", smm_last_error(ctx)); return 3; }
  double s = 0, sy = 0;
  for (size_t i = 0; i < per * R; i++) s += us[i];
  for (int i = 0; i < 16; i++) sy += sum_ys[i];
  printf("%.17g %.17g %lld\n", s, sy, (long long)smm_ctx_launch_count(ctx));
  free(us);
  smm_ctx_destroy(ctx);
  return 0;
}
