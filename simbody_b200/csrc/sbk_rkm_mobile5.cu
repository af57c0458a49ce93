// sbk_rkm_mobile5.cu -- integrator kernels (fixed-step task queue + error-controlled) of the thread-per-instance plan
// for the mobilizer set JM_MOBILE5; see sbk_tpi.cuh.
#include "sbk_tpi.cuh"
SBK_DEFINE_RKM_VARIANT(launchTpiRkmMobile5, SBK_HEAVY_MINB, JM_MOBILE5)
