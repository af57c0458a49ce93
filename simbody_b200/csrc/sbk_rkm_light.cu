// sbk_rkm_light.cu -- integrator kernels (fixed-step task queue + error-controlled) of the thread-per-instance plan
// for the mobilizer set JM_LIGHT; see sbk_tpi.cuh.
#include "sbk_tpi.cuh"
SBK_DEFINE_RKM_VARIANT(launchTpiRkmLight, SBK_TPI_MINBLOCKS, JM_LIGHT)
