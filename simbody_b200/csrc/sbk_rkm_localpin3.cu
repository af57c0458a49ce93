// Pin-only body-frame integrator at twelve warps per SM (168 registers: no spills for this mobilizer kind): the fixed-step
// kernel as one 384-thread CTA of three 128-instance work groups over one staged copy of the tables, 17-row prefetch slots
// (what a Pin needs), 67 work rows per instance.
#define SBK_LPF_ROWS 17
#include "sbk_tpi.cuh"
SBK_DEFINE_RKM_VARIANT_VC(launchTpiRkmLocalPin_m3, 3, JM_PIN | JM_LOCAL)
