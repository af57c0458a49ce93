// sbk_ctree.cu -- plan 5: cluster-level-parallel fixed-step integrator for wide trees in small batches.
//
// One thread-block CLUSTER owns 32 instances (lanes of every warp = those instances).  The tree is cut at the first level with at
// least one body per warp of the cluster: every body of that level roots a subtree that ONE warp walks depth-first with no barrier
// at all (links in the warp's carry, like the thread-per-instance plan); the few levels above the cut run level-parallel on the
// cluster's warps with the cluster's hardware barrier (barrier.cluster) between levels (or, selectable, on the first CTA's eight
// warps with __syncthreads) -- no grid-wide barrier, clusters never wait for each other.  Body steps, prefetch and the step logic are those of the fused body-frame integrator
// (sbk_local.cuh, sbk_lrkm.cuh) in level order (sbk_ltree.cuh).  Replaces the grid-level plan's 141 grid barriers per step
// (ncu, round 1: 46% of the stall samples) by 100 cluster barriers over a 2.5x shorter instruction stream per body.
#define SBK_CARRY_STRIDE_DEVICE_THREADS 256
#include <algorithm>
#include <cooperative_groups.h>
#include "sbk_kernels.cuh"
#include "sbk_ltree.cuh"

namespace cg = cooperative_groups;

namespace sbkd {
namespace {

constexpr int CT_THREADS = 256;

__device__ __forceinline__ void fillCtxTree(Ctx& c, const KArgs& a) {
    c.bodies = nullptr; c.children = nullptr; c.forces = nullptr;
    c.nb = a.nb; c.nq = a.nq; c.nu = a.nu; c.nquat = a.nquat;
    c.gx = a.gx; c.gy = a.gy; c.gz = a.gz;
    c.cache = a.cache; c.cStride = a.cStride; c.cInstStride = a.cInstStride; c.cSpan = a.cSpan; c.cShift = a.cShift; c.cMask = a.cMask;
    c.sStride = a.N; c.sInstStride = 1; c.sSpan = (long long)(a.nq + a.nu)*BLK_LANES;
    c.q = a.yb; c.u = a.yb + (long long)a.nq*BLK_LANES;
    c.qdot = nullptr; c.udot = nullptr; c.qdotdot = nullptr; c.qerr = nullptr;
    c.fmobIn = nullptr; c.FbodyIn = nullptr; c.fmobOut = nullptr; c.FbodyOut = nullptr; c.vecIn = nullptr; c.vecOut = nullptr;
    c.status = a.status;
    c.tp = nullptr; c.ntp = 0; c.f2 = nullptr;
}

template <int JMASK>
__global__ void __launch_bounds__(CT_THREADS, 1) ctreeRkmKernel(const KArgs a, const int CS, const int K) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    // K clusters share one group of 32 instances (K > 1 only when all the grid's clusters are resident at once: they wait for
    // each other): cluster kc of the group holds warps kc*CS*8 .. of the group's K*CS*8
    const int cid = blockIdx.x / CS, iw = cid / K, kc = cid - iw*K;
    const int lane = threadIdx.x & 31, wc = (kc*CS + crank)*(CT_THREADS/32) + (threadIdx.x >> 5), nw = K*CS*(CT_THREADS/32);
    const int inst = iw*32 + lane; const bool active = inst < a.N;
    Ctx c; fillCtxTree(c, a);
    LTables T; T.bodies = reinterpret_cast<const LBody*>(a.ltablesLevel); T.children = reinterpret_cast<const int*>(a.ltablesLevel + a.lchildrenOff);
    T.forces = reinterpret_cast<const ForceConst*>(a.ltablesLevel + a.lforcesOff); T.fcoef = reinterpret_cast<const double*>(a.ltablesLevel + a.lfcoefOff);
    // this warp's task lists (host-built, topology.cpp: cutTreeForWarps): subtree walks below the cut level, the levels above it on the
    // eight warps of the cluster's first CTA (__syncthreads between levels), one cluster barrier per sweep between the two parts
    const int* lists = reinterpret_cast<const int*>(a.ltablesLevel + a.llistsOff);
    const int* listStart = reinterpret_cast<const int*>(a.ltablesLevel + a.llistStartOff);
    const int* lstIn = lists + listStart[wc]; const int* lstOut = lists + listStart[nw + wc];
    double* cy = reinterpret_cast<double*>(smem) + threadIdx.x;
    LBodySlots B; B.slot = reinterpret_cast<LBody*>(smem + (size_t)LFCARRY_ROWS*CT_THREADS*sizeof(double)) + (threadIdx.x >> 5)*LT_BODY_SLOTS;
    LRkmWork w; w.Y = a.yb; w.W = a.ys; w.F0 = a.f0; w.F2 = a.fa; w.F3 = a.fb; w.Ynext = a.yb;
    w.accuracy = a.accuracy; w.consTol = a.consTol; w.useInfNorm = a.useInfNorm; w.projectEveryStep = a.projectEveryStep;
    const int ny = a.nq + a.nu;
    // barrier.cluster.arrive (release) / wait (acquire) order the global-memory hand-over rows between the warps of the cluster;
    // every such row is written with st.cg and read with ld.cg (L2), never through a possibly stale L1 line
    // per-group scratch in global memory: partial error sums [nw][3][32], the group-uniform "some instance projected" flag and the
    // arrival counter of the barrier between the group's clusters (zeroed by the launcher)
    double* part = a.treeScratch + (size_t)iw*((size_t)nw*3*32 + 32);
    int* flag = reinterpret_cast<int*>(part + (size_t)nw*3*32);
    unsigned* xcount = reinterpret_cast<unsigned*>(flag) + 4;
    unsigned xphase = 0;
    // barrier of the group's K clusters: every thread publishes its global writes (membar.gl), the cluster meets, one thread per
    // cluster arrives on the counter (release) and waits for the K arrivals of this phase (acquire), the cluster meets again
    auto crossSync = [&]() {
        __threadfence();
        cluster.sync();
        ++xphase;
        if (crank == 0 && threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(xcount) : "memory");
            unsigned seen;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(xcount) : "memory"); } while (seen < xphase*(unsigned)K);
        }
        cluster.sync();
    };
    auto groupSync = [&](const bool cross) { if (cross && K > 1) crossSync(); else cluster.sync(); };
    auto allSync = [&]() { if (K > 1) crossSync(); else { __threadfence(); cluster.sync(); } };
    auto topSync = [&]() { __syncthreads(); };
    auto reduce = [&](double& q, double& u, double& qt) {
        __stcg(part + ((size_t)wc*3 + 0)*32 + lane, q); __stcg(part + ((size_t)wc*3 + 1)*32 + lane, u); __stcg(part + ((size_t)wc*3 + 2)*32 + lane, qt);
        allSync();
        if (wc == 0) {                              // fixed summation order: deterministic
            double s0 = 0, s1 = 0, s2 = 0;
            for (int k = 0; k < nw; ++k) {
                const double p0 = __ldcg(part + ((size_t)k*3 + 0)*32 + lane), p1 = __ldcg(part + ((size_t)k*3 + 1)*32 + lane), p2 = __ldcg(part + ((size_t)k*3 + 2)*32 + lane);
                if (a.useInfNorm) { s0 = normMax(s0, p0); s1 = normMax(s1, p1); s2 = normMax(s2, p2); } else { s0 += p0; s1 += p1; s2 += p2; }
            }
            q = s0; u = s1; qt = s2;
        }
    };
    // state into the blocked layout, Ground's link rows
    if (active) for (int i = wc; i < ny; i += nw) a.yb[stateIndex<true>(c, inst, i)] = __ldcg(a.y + (long long)i*a.N + inst);
    if (active && wc == 0) lLevelGround(c, T, inst);
    allSync();
    int vb = 0, par = 0; bool velValid = false;
    double err = 0; int nproj = 0;
#pragma unroll 1
    for (int s = 0; s < a.nsteps; ++s) {
        const RkmStepResult r = lListStep<JMASK>(c, T, lstIn, lstOut, B, inst, active, lane, cy, w, a.h, vb, velValid, wc, par, topSync, groupSync, reduce);
        if (wc == 0) {
            const unsigned any = __ballot_sync(0xffffffffu, active && r.projected);
            if (lane == 0) __stcg(flag, any != 0 ? 1 : 0);
            if (active) { err = r.errNorm; nproj += r.projected; if (a.status && !finiteNorm(r.errNorm)) atomicOr(a.status + inst, 1); }
        }
        allSync();
        velValid = __ldcg(flag) == 0;
    }
    lpfWaitAll();
    if (active) for (int i = wc; i < ny; i += nw) __stcg(a.y + (long long)i*a.N + inst, a.yb[stateIndex<true>(c, inst, i)]);
    if (active && wc == 0) { a.tcur[inst] += a.nsteps*a.h; a.errNorm[inst] = err; a.projCount[inst] += nproj; }
}

} // namespace

// scratch doubles per instance warp for a cluster of CS CTAs
size_t ctreeScratchDoubles(int N, int CS) { return (size_t)((N + 31)/32)*((size_t)CS*(CT_THREADS/32)*3*32 + 32); }

static cudaError_t ctreeConfig(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int nclusters, int CS, cudaStream_t stream, bool cooperative = false) {
    auto kernel = ctreeRkmKernel<JM_MOBILE5 | JM_LOCAL>;
    const size_t smemBytes = (size_t)LFCARRY_ROWS*CT_THREADS*sizeof(double) + (size_t)(CT_THREADS/32)*LT_BODY_SLOTS*sizeof(LBody);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
    if (e != cudaSuccess) return e;
    if (CS > 8) { e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); if (e != cudaSuccess) return e; }
    cfg = cudaLaunchConfig_t();
    cfg.gridDim = dim3((unsigned)(nclusters*CS)); cfg.blockDim = dim3(CT_THREADS); cfg.dynamicSmemBytes = smemBytes; cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // clusters that wait for each other: a cooperative launch starts only when the whole grid is resident
    if (cooperative) { attr[1].id = cudaLaunchAttributeCooperative; attr[1].val.cooperative = 1; cfg.numAttrs = 2; }
    return cudaSuccess;
}
// How many clusters of CS CTAs of the integrator kernel the device holds at once (0 on error).
int ctreeMaxActiveClusters(int CS) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
    if (ctreeConfig(cfg, attr, 1, CS, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ctreeRkmKernel<JM_MOBILE5 | JM_LOCAL>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// K clusters per group of 32 instances (the task lists in a.ltablesLevel are cut for K*CS*8 warps).  K > 1 needs every cluster
// of the grid resident at once (ctreeMaxActiveClusters): the caller checks.
cudaError_t launchCtreeRkm(const KArgs& a, int CS, int K, cudaStream_t stream) {
    const int groups = (a.N + 31)/32;
    cudaError_t e = cudaMemsetAsync(a.treeScratch, 0, ctreeScratchDoubles(a.N, CS*K)*sizeof(double), stream);   // barrier counters
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
    e = ctreeConfig(cfg, attr, groups*K, CS, stream, K > 1);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, ctreeRkmKernel<JM_MOBILE5 | JM_LOCAL>, a, CS, K);
}
// Largest portable cluster size (8, 4, 2, 1) of which at least one cluster can be resident on this device.
int ctreeMaxClusterSize() {
    auto kernel = ctreeRkmKernel<JM_MOBILE5 | JM_LOCAL>;
    const size_t smemBytes = (size_t)LFCARRY_ROWS*CT_THREADS*sizeof(double) + (size_t)(CT_THREADS/32)*LT_BODY_SLOTS*sizeof(LBody);
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes) != cudaSuccess) { cudaGetLastError(); return 0; }
    // (16-CTA clusters are legal on B200 but measured 2x slower here: fewer of them are co-resident)
    for (int CS = 8; CS >= 1; CS /= 2) {
        if (CS > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)CS); cfg.blockDim = dim3(CT_THREADS); cfg.dynamicSmemBytes = smemBytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n >= 1) return CS;
        cudaGetLastError();
    }
    return 0;
}

} // namespace sbkd
