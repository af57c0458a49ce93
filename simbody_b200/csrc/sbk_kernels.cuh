// sbk_kernels.cuh -- kernel argument block and launch wrappers (implemented in sbk_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sbk_rkm.cuh"

namespace sbkd {

// Batch-shared tables live in ONE contiguous device blob so that a CTA stages them into shared
// memory with a single TMA bulk copy (cp.async.bulk + mbarrier):
//   [ BodyConst[nb] | int children[...] (16B padded) | ForceConst forces[...] (16B padded) ]
struct KArgs {
    const unsigned char* tables;     // device blob
    uint32_t tableBytes, childrenOff, forcesOff, stageInSmem;
    uint32_t levelOrderOff, levelStartOff; int nlevels, plan;
    uint32_t tpOff; int ntp; double* f2;     // two-point force elements (TwoPointConst[ntp] in the tables blob) and their body forces [nb*6][N]
    // body-frame integrator path (sbk_local.cuh): [ LBody[nb] | children | forces ], null when the model has other mobilizers
    const unsigned char* ltablesLevel;   // the same blob with level-order link flags (plan 5, sbk_ltree.cuh)
    double* treeScratch;                 // plan 5: per-cluster partial error sums + flags
    uint32_t llistsOff, llistStartOff; int nsub, cutLevel;    // plan 5: per-warp task lists (in the ltablesLevel blob)
    const unsigned char* ltables; uint32_t ltableBytes, lchildrenOff, lforcesOff, lstageInSmem, lfcoefOff, lpad_;
    int localMinB, jointMask;                  // localMinB: register budget variant of the body-frame integrator kernels (2, 3, 4 CTAs / SM)                      // bit JT_x set = mobilizer kind x present in the model
    long long cStride, cInstStride, cSpan;  // cache addressing: base_b + field*cStride + (inst>>cShift)*cSpan + (inst&cMask)*cInstStride
    int cShift, cMask;
    int nb, nq, nu, nquat;
    double gx, gy, gz;
    int N;
    double* cache;
    double* y;          // state [nq+nu][N]
    double* yb;         // CTA-blocked working copy of y for the thread-per-instance integrator kernels
    double* ydot;       // [nq+nu][N] (qdot, udot)
    double* qdotdot;    // [nq][N]
    double* qerr;       // [nquat][N]
    const double* fmobIn; const double* FbodyIn;
    double* fmobOut; double* FbodyOut;
    const double* vecIn; double* vecOut;
    int* status;
    // RKM
    double* y0; double* f0; double* fa; double* fb; double* ys;
    double* tcur;       // [N] time
    double h; int nsteps;
    double accuracy, consTol; int useInfNorm, projectEveryStep;
    double* errNorm;    // [N]
    int* projCount;     // [N] accumulated
    // error-controlled stepping (OP_RKM_ADAPT)
    double tFinal, minStep, maxStep; int allowInterp, maxAttempts;
    double* hcur;       // [N] current step size
    double* lastStep;   // [N] size of the last accepted step
    int* stepsTaken;    // [N] accumulated
    int* attempts;      // [N] accumulated
    int roundSync;      // fixed-step integrator: the CTAs take their tasks round-robin and meet at a grid barrier before every round (set by launchOp)
    int* taskCounter;   // fixed-step integrator task queue: a 64-bit next-task counter (also plan 4's barrier counter); blockDone = taskCounter + 2
    int* blockDone;     // steps completed per block of 128 instances (this launch)
    int* lflags;        // [N] fused body-frame integrator: 1 = the velocity data left by the previous step are current
};

enum KernelOp {
    OP_KIN = 0,        // sweeps A+B
    OP_ABI,            // sweep C (+ZB)
    OP_EVAL,           // full realize(Acceleration)
    OP_CALCACC,        // sweeps D+E with caller forces
    OP_MULM, OP_MULMINV, OP_RESID,
    OP_RKM,
    OP_RKM_ADAPT       // error-controlled RKM to a final time
};

// Launch one operation for the thread-per-instance plan on `stream`.
cudaError_t launchTpi(KernelOp op, const KArgs& a, cudaStream_t stream);
// Integrator kernels of the thread-per-instance plan, one family per set of mobilizer kinds (op = OP_RKM / OP_RKM_ADAPT).
cudaError_t launchTpiRkmAll(KernelOp op, const KArgs& a, cudaStream_t stream);      // ground-frame integrator, every supported mobilizer
// body-frame sweeps (Pin / Slider / Universal / Ball / Free, and Pin only), built for 2 resident CTAs per SM (255 registers; the
// 3- and 4-CTA builds of sbk_rkm_local.inc measured no gain: shared memory, not registers, limits the occupancy)
cudaError_t launchTpiRkmLocal_m2(KernelOp op, const KArgs& a, cudaStream_t stream); cudaError_t launchTpiRkmLocalPin_m2(KernelOp op, const KArgs& a, cudaStream_t stream);
cudaError_t launchTpiRkmLocalPin_m3(KernelOp op, const KArgs& a, cudaStream_t stream); size_t launchTpiRkmLocalPin_m3_workBytes();
// Register-resident fused plan (serial chains of 1-2 Pin/Slider[/Universal] mobilizers).
bool fusedPlanSupports(int nb, const int* joints /*[nb]*/);
cudaError_t launchFusedRkm(const KArgs& a, const int* joints, bool adaptive, cudaStream_t stream);
// Level-parallel plan (wide trees, small batches): one CTA per instance, threads over the bodies
// of a tree level, __syncthreads between levels.
cudaError_t launchLp(KernelOp op, const KArgs& a, cudaStream_t stream);
// Grid-level-parallel fixed-step integrator (plan 4): persistent cooperative grid, work items = (body of a level, warp of
// instances), grid barriers between levels.  Uses the thread-per-instance record layout and the [slot][N] work vectors.
cudaError_t launchGlRkm(const KArgs& a, cudaStream_t stream);
// Every operation of plan 4 (OP_RKM and the API operations; OP_RKM_ADAPT is not available).
cudaError_t launchGl(KernelOp op, const KArgs& a, cudaStream_t stream);
// Integrator::initialize's forced projection: normalise every quaternion of y, count one projection per instance.
cudaError_t launchInitProject(const KArgs& a, cudaStream_t stream);
// Plan 5 (cluster-level-parallel, sbk_ctree.cu): fixed-step integrator; CS = CTAs per cluster (one cluster per 32 instances).
cudaError_t launchCtreeRkm(const KArgs& a, int CS, int K, cudaStream_t stream);
int ctreeMaxActiveClusters(int CS);
int ctreeMaxClusterSize();
size_t ctreeScratchDoubles(int N, int CS);
// Ground record (identity transform, zero velocity/acceleration) for every instance.
cudaError_t launchInitGround(const KArgs& a, cudaStream_t stream);
// dst[k*len + i] <-> src[i*N + k]
cudaError_t launchTranspose(const double* src, double* dst, int rows, int cols, cudaStream_t stream);
// Gather one per-body cache field (width doubles at field offset) into out[(b*width+i)*N + k].
cudaError_t launchGatherBodyField(const KArgs& a, int fieldOffset, int width, double* out, cudaStream_t stream);
// Kinetic / potential energy per instance from the realized records (ke, pe: device [N], nullable).
cudaError_t launchEnergy(const KArgs& a, double* ke, double* pe, cudaStream_t stream);
// Mobilizer reaction forces at the M frame origins, in Ground: out [nb*6][N] (needs a realized acceleration stage).
cudaError_t launchReaction(const KArgs& a, double* out, cudaStream_t stream);
// System Jacobian products from the realized position records: Jv [nb*6][N] = J v;  JtF [nu][N] = ~J F (z: scratch [nb*6][N]).
cudaError_t launchJacobian(const KArgs& a, const double* v, double* out, cudaStream_t stream);
cudaError_t launchJacobianTranspose(const KArgs& a, const double* F, double* z, double* out, cudaStream_t stream);
// Composite body inertias from the realized position records: out [nb*10][N] (mass, com, unit inertia).
cudaError_t launchCompositeBodyInertias(const KArgs& a, double* out, cudaStream_t stream);
// Memory-pattern probe for the thread-per-instance record layout (diagnostics).
cudaError_t launchMemPattern(double* buf, int N, int nb, int rowsIn, int rowsOut, int sweeps, int minBlocks, cudaStream_t stream);
// FP64 FMA throughput probe: returns flops executed; used by bench.py to measure the FP64 roofline.
cudaError_t launchDfmaProbe(double* out, int iters, int blocks, int threads, cudaStream_t stream);

} // namespace sbkd
