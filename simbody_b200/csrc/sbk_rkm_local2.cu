// Thread-per-instance integrator kernels, body-frame sweeps, 2 resident CTAs per SM; see sbk_rkm_local.inc
#define SBK_LOCAL_MINB 2
#include "sbk_rkm_local.inc"
