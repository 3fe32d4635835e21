// temporary: PDE entry points until pde.cu lands
#include "smm_internal.h"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { return smm_fail(nullptr, SMM_ERR_UNSUPPORTED, #name ": not built yet"); }
STUB(smm_pde_rhs, smm_ctx*, const smm_grid*, const smm_params*, int32_t, int32_t, int32_t, const double*, const double*, const double*, const double*, const double*, double*, double*, double*)
STUB(smm_pde_solve, smm_ctx*, const smm_grid*, const smm_params*, int32_t, int32_t, int32_t, const double*, const double*, const double*, const double*, const double*, double, int32_t, int32_t, int32_t, const int32_t*, double*, double*, double*, double*, double*, double*)
STUB(smm_pde_batch_create, smm_ctx*, const smm_grid*, const smm_params*, int32_t, int32_t, int32_t, const double*, const double*, const double*, const double*, const double*, int32_t, smm_pde_batch**)
STUB(smm_pde_batch_step, smm_ctx*, smm_pde_batch*, double, int32_t, int32_t)
STUB(smm_pde_batch_read, smm_ctx*, smm_pde_batch*, double*, double*, double*)
STUB(smm_pde_batch_sum_dev, smm_ctx*, smm_pde_batch*, double*, double*, double*)
STUB(smm_pde_batch_destroy, smm_ctx*, smm_pde_batch*)
