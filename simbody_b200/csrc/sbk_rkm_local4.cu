// Thread-per-instance integrator kernels, body-frame sweeps, 4 resident CTAs per SM; see sbk_rkm_local.inc
#define SBK_LOCAL_MINB 4
#include "sbk_rkm_local.inc"
