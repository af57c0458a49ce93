// sbk_rkm_all.cu -- integrator kernels (fixed-step task queue + error-controlled) of the thread-per-instance plan
// for the mobilizer set JM_ALL; see sbk_tpi.cuh.
#include "sbk_tpi.cuh"
SBK_DEFINE_RKM_VARIANT(launchTpiRkmAll, SBK_HEAVY_MINB, JM_ALL)
