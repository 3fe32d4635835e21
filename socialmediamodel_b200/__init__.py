"""socialmediamodel_b200 -- B200-native (sm_100a) hot path of LuzieH/SocialMediaModel.

Exports the reference's API for the hot path (src/SocialMediaModel.jl:12-16): the parameter
constructors of setting.jl, ABMsolve, ABMsolveensemble, PDEsolve, PDEsolveensemble and the PDE
right-hand side ``f``.  Everything computes in libsmm_b200.so (hand-written CUDA, C ABI in
include/smm_b200.h); there is no CPU fallback.
"""
from .setting import (ABMconstruct, PDEconstruct, PDEconstructcoarse, PDEconstructmeso, Struct, parameters,
                      parameterscontrol, parametersstronginf)
from .abm import ABMsolve, abm_draws, abm_init, abm_solve_raw
from .ensemble import (ABMsolveensemble, PDEsolveensemble, abm_ensemble_raw, ensemble_means, kde, runensembles,
                       savetimes_abm)
from .pde import (ArrayPartition, PDEBatch, PDESolution, PDEsolve, constructinitial, default_dt, f, gaussian, pde_rhs,
                  pde_solve_adaptive_raw, pde_solve_raw, randominit, sol2uyz, uniforminit)
from .control import PDEsolvefixedtargets, PDEsolvefixedtargetsfast, PDEsolveopt, fixedtargets_objectives, prep
from .population import ABMPopulation
from ._lib import Context, SmmError

__all__ = ["parameters", "parametersstronginf", "ABMconstruct", "PDEconstruct", "PDEconstructcoarse",
           "PDEconstructmeso", "parameterscontrol", "ABMsolve", "ABMsolveensemble", "PDEsolve", "PDEsolveensemble", "f", "sol2uyz", "Context", "SmmError"]
