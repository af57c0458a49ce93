// sbk_rkm_pin.cu -- integrator kernels (fixed-step task queue + error-controlled) of the thread-per-instance plan
// for the mobilizer set JM_PIN; see sbk_tpi.cuh.
#include "sbk_tpi.cuh"
SBK_DEFINE_RKM_VARIANT(launchTpiRkmPin, SBK_TPI_MINBLOCKS, JM_PIN)
