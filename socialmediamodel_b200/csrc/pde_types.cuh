// pde_types.cuh -- per-instance device structures of the PDE batch, shared by pde.cu and pde_cluster.cu
#pragma once

namespace {

constexpr int LMAX = 16;

struct PdeInst {  // per-instance constants
  double a, b, c, eta, D, frictionI, frictionM, speed1, speed2, tgt1[2], tgt2[2];
  int cinf[LMAX];
  int cmed[2];
};

struct PdeCoef {  // per-instance stage data: written by the finalising CTA, read by the next stage
  double y[LMAX][2], z[2][2];    // stage values of the ODE unknowns
  double dy[LMAX][2], dz[2][2];  // their right-hand sides at this stage (pde.jl:239-265)
  double R[2][LMAX];             // eta * r((m_ij - m_i'j)/(m_ij + m_i'j))  (pde.jl:43)
};

struct PdeOde {  // RK bookkeeping of the ODE unknowns
  double yn[LMAX][2], zn[2][2], yacc[LMAX][2], zacc[2][2];
  double yerr[LMAX][2], zerr[2][2];  // embedded error estimate (adaptive stepping)
};

}  // namespace
