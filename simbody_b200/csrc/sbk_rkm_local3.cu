// Thread-per-instance integrator kernels, body-frame sweeps, 3 resident CTAs per SM; see sbk_rkm_local.inc
#define SBK_LOCAL_MINB 3
#include "sbk_rkm_local.inc"
