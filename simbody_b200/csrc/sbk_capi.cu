// sbk_capi.cu -- implementation of the C ABI in include/sbk.h (host side, CUDA runtime).
//
// Host responsibilities: topology compilation (topology.cpp), the structure-of-arrays batched
// State in HBM, stage bookkeeping that mirrors the reference's Stage checks, and kernel
// launches on the batch's stream.  There is no CPU compute path: without a usable CUDA device
// every compute entry point returns SBK_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "sbk.h"
#include "topology.h"
#include "sbk_kernels.cuh"

using namespace sbkd;

namespace {
thread_local std::string g_lastError;
int fail(int code, const std::string& msg) { g_lastError = msg; return code; }
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return fail(SBK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

enum Stage { ST_EMPTY = 0, ST_POSITION = 1, ST_VELOCITY = 2, ST_ACCELERATION = 3 };
} // namespace

struct sbk_batch {
    const sbk_topology* topo = nullptr;
    int N = 0, device = 0, plan = 1;
    cudaStream_t stream = nullptr; bool ownStream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    KArgs a;                      // device pointers + constants
    unsigned char* dTables = nullptr; unsigned char* dLTables = nullptr; unsigned char* dLTablesLevel = nullptr;
    int clusterSize = 0;          // plan 5: CTAs per cluster
    int clustersPerGroup = 1;     // plan 5: clusters that share one group of 32 instances
    double *dOpA = nullptr, *dOpB = nullptr, *dOpF = nullptr, *dOpOut = nullptr, *dScratch = nullptr;
    size_t scratchDoubles = 0;
    int stage = ST_EMPTY; bool abiValid = false, accelValid = false;
    int64_t launches = 0, stepsTaken = 0, realizations = 0;
    bool adaptiveInit = false;
    int64_t adaptStepsSeen = 0, adaptAttemptsSeen = 0;   // sums of the per-instance adaptive counters already folded into the stats
    bool integInit = false;       // Integrator::initialize's forced projection has run for the current state
    double lastKernelMs = 0;
    long long recTotal = 0;
};

namespace {

int useDevice(const sbk_batch* b) {
    cudaError_t e = cudaSetDevice(b->device);
    if (e != cudaSuccess) return fail(SBK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return SBK_OK;
}
int launch(sbk_batch* b, KernelOp op) {
    CUDA_TRY(b->plan == 3 ? launchLp(op, b->a, b->stream) : (b->plan == 4 || b->plan == 5) ? launchGl(op, b->a, b->stream) : launchTpi(op, b->a, b->stream));
    b->launches++;
    return SBK_OK;
}
int ensureScratch(sbk_batch* b, size_t doubles) {
    if (doubles <= b->scratchDoubles) return SBK_OK;
    if (b->dScratch) cudaFree(b->dScratch);
    b->dScratch = nullptr; b->scratchDoubles = 0;
    CUDA_TRY(cudaMalloc(&b->dScratch, doubles*sizeof(double)));
    b->scratchDoubles = doubles;
    return SBK_OK;
}
// host SoA [rows][N] -> device
int h2d(sbk_batch* b, double* dst, const double* src, size_t rows) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, rows*b->N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
    return SBK_OK;
}
int d2h(sbk_batch* b, double* dst, const double* src, size_t rows) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, rows*b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}
int needStage(const sbk_batch* b, int st, const char* who) {
    if (b->stage < st) {
        static const char* names[] = {"Empty", "Position", "Velocity", "Acceleration"};
        return fail(SBK_ERR_STAGE, std::string(who) + ": state must be realized to Stage::" + names[st] +
                                   " (current: " + names[b->stage] + ")");
    }
    return SBK_OK;
}
} // namespace

extern "C" {

int sbk_version(void) { return SBK_VERSION; }
const char* sbk_last_error(void) { return g_lastError.c_str(); }

int sbk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

// ---- topology -----------------------------------------------------------------------------
sbk_topology* sbk_topology_create(const sbk_body_desc* bodies, int nb, const sbk_force_desc* forces, int nf) {
    return sbk_topology_create_ex(bodies, nb, forces, nf, 0u);
}
sbk_topology* sbk_topology_create_ex(const sbk_body_desc* bodies, int nb, const sbk_force_desc* forces, int nf, unsigned flags) {
    if (!bodies || nb < 1 || nf < 0 || (nf > 0 && !forces)) { fail(SBK_ERR_ARG, "sbk_topology_create: bad arguments"); return nullptr; }
    try {
        sbk::ModelSpec spec; spec.name = "user"; spec.useEulerAngles = (flags & SBK_TOPOLOGY_EULER_ANGLES) != 0;
        spec.bodies.assign(bodies, bodies + nb);
        if (nf) spec.forces.assign(forces, forces + nf);
        sbk_topology* t = new sbk_topology();
        try { sbk::compileTopology(spec, *t); } catch (...) { delete t; throw; }
        return t;
    } catch (const std::exception& e) { fail(SBK_ERR_TOPOLOGY, e.what()); return nullptr; }
}
sbk_topology* sbk_topology_from_text(const char* text) {
    if (!text) { fail(SBK_ERR_ARG, "sbk_topology_from_text: null text"); return nullptr; }
    try {
        sbk::ModelSpec spec = sbk::fromText(text);
        sbk_topology* t = new sbk_topology();
        try { sbk::compileTopology(spec, *t); } catch (...) { delete t; throw; }
        return t;
    } catch (const std::exception& e) { fail(SBK_ERR_TOPOLOGY, e.what()); return nullptr; }
}
void sbk_topology_destroy(sbk_topology* t) { delete t; }
int sbk_topology_counts(const sbk_topology* t, int* nb, int* nq, int* nu, int* nquat, int* nlevels) {
    if (!t) return fail(SBK_ERR_ARG, "sbk_topology_counts: null topology");
    if (nb) *nb = t->nb; if (nq) *nq = t->nq; if (nu) *nu = t->nu; if (nquat) *nquat = t->nquat; if (nlevels) *nlevels = t->nlevels;
    return SBK_OK;
}
int sbk_topology_slots(const sbk_topology* t, int* q0, int* nq, int* u0, int* nu, int* level) {
    if (!t) return fail(SBK_ERR_ARG, "sbk_topology_slots: null topology");
    for (int b = 0; b < t->nb; ++b) {
        if (q0) q0[b] = t->q0[b]; if (nq) nq[b] = t->nqOf[b]; if (u0) u0[b] = t->u0[b]; if (nu) nu[b] = t->nuOf[b];
        if (level) level[b] = t->level[b];
    }
    return SBK_OK;
}
int sbk_model_text(const char* name, int n, char* buf, int cap) {
    if (!name) { fail(SBK_ERR_ARG, "sbk_model_text: null name"); return -1; }
    try {
        const std::string s = sbk::toText(sbk::makeNamedModel(name, n));
        if (buf && (int)s.size() + 1 <= cap) std::memcpy(buf, s.c_str(), s.size() + 1);
        return (int)s.size() + 1;
    } catch (const std::exception& e) { fail(SBK_ERR_ARG, e.what()); return -1; }
}

// ---- batch --------------------------------------------------------------------------------
static bool fusedOk(const sbk_batch* b) {
    if (!b->topo->twoPoint.empty()) return false;
    std::vector<int> joints(b->topo->nb);
    for (int i = 0; i < b->topo->nb; ++i) joints[i] = b->topo->bodies[i].joint;
    return b->topo->isChain && fusedPlanSupports(b->topo->nb, joints.data());
}
static int autoPlan(const sbk_batch* b) {
    if (fusedOk(b)) return 2;
    // wide tree + batch too small to fill 148 SMs with one thread per instance: level-parallel plans (5 = clusters on the
    // body-frame sweeps where the model allows and the device can host a cluster, else 4 = cooperative grid)
    if (b->N < 16384 && b->topo->nb >= 128 && b->topo->maxLevelWidth >= 32) {
        const char* e = getenv("SBK_NOLOCAL");
        if (b->topo->localOk && !(e && atoi(e)) && ctreeMaxClusterSize() >= 2) return 5;
        return 4;
    }
    return 1;
}
// (Re)build the batch-shared tables and the per-body cache for an execution plan.
//   plans 1/2/4: CTA-blocked records, see below (plan 4 differs from 1 only in its fixed-step integrator kernel)
//   plan 3   : cache[i*(KMAX*nb) + k*nb + pos_b], pos_b = position in (level, joint) order, so
//              that the threads of a CTA (bodies of one level) touch adjacent addresses.
static int configurePlan(sbk_batch* b, int plan) {
    const sbk_topology* t = b->topo; const int n = b->N; KArgs& a = b->a;
    std::vector<BodyConst> bodies = t->bodies;
    std::vector<int> order = t->levelOrder;
    long long cacheDoubles = 0;
    if (plan == 3) {
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            if (t->level[x] != t->level[y]) return t->level[x] < t->level[y];
            return t->bodies[x].joint < t->bodies[y].joint; });
        std::vector<int> pos(t->nb);
        for (int i = 0; i < t->nb; ++i) pos[order[i]] = i;
        for (int i = 0; i < t->nb; ++i) bodies[i].cacheBase = pos[i];
        a.cStride = t->nb; a.cInstStride = 0; a.cSpan = (long long)CACHE_RECORD_MAX*t->nb; a.cShift = 0; a.cMask = 0;
        cacheDoubles = a.cSpan*n;
    } else {
        // CTA-blocked records: [block of 128 instances][record field][lane]; the integrator kernels use the
        // row stride 128 as a compile-time constant, the API kernels read it from the context.
        long long off = 0;
        for (int i = 0; i < t->nb; ++i) { bodies[i].cacheBase = off*BLK_LANES; off += (i == 0) ? F_H : cacheRecordSize(t->nuOf[i]); }
        a.cStride = BLK_LANES; a.cInstStride = 1; a.cSpan = off*BLK_LANES; a.cShift = 7; a.cMask = BLK_LANES - 1;
        cacheDoubles = off*BLK_LANES*(((long long)n + BLK_LANES - 1)/BLK_LANES);
        b->recTotal = off;
    }
    for (int i = 0; i < t->nb; ++i) bodies[i].parentCacheBase = bodies[bodies[i].parent].cacheBase;
    auto pad16 = [](size_t x) { return (x + 15)/16*16; };
    const size_t bodiesBytes = pad16(bodies.size()*sizeof(BodyConst));
    const size_t childBytes  = pad16(t->children.size()*sizeof(int));
    const size_t forceBytes  = pad16(t->forces.size()*sizeof(ForceConst));
    const size_t orderBytes  = pad16(order.size()*sizeof(int));
    const size_t startBytes  = pad16(t->levelStart.size()*sizeof(int));
    std::vector<TwoPointConst> tps = t->twoPoint;
    for (TwoPointConst& tp : tps) { tp.cacheBase1 = bodies[tp.body1].cacheBase; tp.cacheBase2 = bodies[tp.body2].cacheBase; }
    const size_t tpBytes = pad16(std::max<size_t>(tps.size(), 1)*sizeof(TwoPointConst));
    std::vector<unsigned char> blob(bodiesBytes + childBytes + forceBytes + orderBytes + startBytes + tpBytes, 0);
    if (!tps.empty()) std::memcpy(blob.data() + bodiesBytes + childBytes + forceBytes + orderBytes + startBytes, tps.data(), tps.size()*sizeof(TwoPointConst));
    a.tpOff = (uint32_t)(bodiesBytes + childBytes + forceBytes + orderBytes + startBytes); a.ntp = (int)tps.size();
    std::memcpy(blob.data(), bodies.data(), bodies.size()*sizeof(BodyConst));
    std::memcpy(blob.data() + bodiesBytes, t->children.data(), t->children.size()*sizeof(int));
    std::memcpy(blob.data() + bodiesBytes + childBytes, t->forces.data(), t->forces.size()*sizeof(ForceConst));
    std::memcpy(blob.data() + bodiesBytes + childBytes + forceBytes, order.data(), order.size()*sizeof(int));
    std::memcpy(blob.data() + bodiesBytes + childBytes + forceBytes + orderBytes, t->levelStart.data(), t->levelStart.size()*sizeof(int));
    a.tableBytes = (uint32_t)blob.size(); a.childrenOff = (uint32_t)bodiesBytes; a.forcesOff = (uint32_t)(bodiesBytes + childBytes);
    a.levelOrderOff = (uint32_t)(bodiesBytes + childBytes + forceBytes); a.levelStartOff = a.levelOrderOff + (uint32_t)orderBytes;
    a.nlevels = t->nlevels; a.plan = plan;
    a.jointMask = 0;
    for (int i = 1; i < t->nb; ++i) a.jointMask |= 1 << t->bodies[i].joint;
    a.stageInSmem = (plan != 3 && blob.size() <= 28*1024) ? 1u : 0u;
    { const char* e = getenv("SBK_NOSTAGE"); if (e && atoi(e)) a.stageInSmem = 0; }   // tuning override: read the tables through L1/L2   // two resident CTAs: 2 x (tables + 84 KB carry) <= 227 KB
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (b->dTables) cudaFree(b->dTables);
    if (a.cache) cudaFree(a.cache);
    b->dTables = nullptr; a.cache = nullptr;
    CUDA_TRY(cudaMalloc(&b->dTables, blob.size()));
    CUDA_TRY(cudaMemcpyAsync(b->dTables, blob.data(), blob.size(), cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    a.tables = b->dTables;
    // body-frame integrator tables (sbk_local.cuh): plans 1 / 4, models made of Pin / Slider / Universal / Ball / Free
    if (b->dLTables) cudaFree(b->dLTables);
    b->dLTables = nullptr; a.ltables = nullptr; a.ltableBytes = 0; a.lstageInSmem = 0;
    if (b->dLTablesLevel) cudaFree(b->dLTablesLevel);
    b->dLTablesLevel = nullptr; a.ltablesLevel = nullptr;
    { const char* e = getenv("SBK_NOLOCAL");
      if (t->localOk && plan != 3 && !(e && atoi(e))) {
        const size_t lb = pad16(t->lbodies.size()*sizeof(LBody));
        const size_t fcoefBytes = pad16(t->lfcoef.size()*sizeof(double));
        std::vector<unsigned char> lblob(lb + childBytes + forceBytes + fcoefBytes, 0);
        std::memcpy(lblob.data() + lb + childBytes + forceBytes, t->lfcoef.data(), t->lfcoef.size()*sizeof(double));
        a.lfcoefOff = (uint32_t)(lb + childBytes + forceBytes);
        a.localMinB = 2;
        std::memcpy(lblob.data(), t->lbodies.data(), t->lbodies.size()*sizeof(LBody));
        std::memcpy(lblob.data() + lb, t->children.data(), t->children.size()*sizeof(int));
        std::memcpy(lblob.data() + lb + childBytes, t->forces.data(), t->forces.size()*sizeof(ForceConst));
        CUDA_TRY(cudaMalloc(&b->dLTables, lblob.size()));
        CUDA_TRY(cudaMemcpyAsync(b->dLTables, lblob.data(), lblob.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        a.ltables = b->dLTables; a.ltableBytes = (uint32_t)lblob.size(); a.lchildrenOff = (uint32_t)lb; a.lforcesOff = (uint32_t)(lb + childBytes);
        a.lstageInSmem = (lblob.size() <= 48*1024 && a.stageInSmem != 0) || (lblob.size() <= 48*1024 && blob.size() > 28*1024) ? 1u : 0u;
        { const char* e2 = getenv("SBK_NOSTAGE"); if (e2 && atoi(e2)) a.lstageInSmem = 0; }
        if (plan == 5) {          // a second blob with the link flags of the cut-tree schedule + the subtree walks, and the clusters' scratch
            b->clusterSize = ctreeMaxClusterSize();
            { const char* e4 = getenv("SBK_CLUSTER"); if (e4 && atoi(e4) > 0) b->clusterSize = std::min(b->clusterSize, atoi(e4)); }   // tuning override
            if (b->clusterSize < 1) return fail(SBK_ERR_CUDA, "plan 5: the device cannot host a thread-block cluster of the integrator kernel");
            // Two clusters per group of 32 instances when the device holds all of them at once (they wait for each other at
            // a barrier in global memory): the subtrees below the cut are half as deep.  C5: 8 groups x 2 clusters x 8 CTAs.
            b->clustersPerGroup = 1;
            { const int groups = (n + 31)/32; int resident = ctreeMaxActiveClusters(b->clusterSize);
              if (b->clusterSize == 8 && !getenv("SBK_CLUSTER")) {
                  // 128 warps per group: 2 clusters of 8 CTAs, else 4 clusters of 4 (a B200 holds 15 clusters of 8 but 36 of 4)
                  if (2*groups <= resident) b->clustersPerGroup = 2;
                  else if (4*groups <= ctreeMaxActiveClusters(4)) { b->clusterSize = 4; b->clustersPerGroup = 4; resident = ctreeMaxActiveClusters(4); }
              }
              const char* e7 = getenv("SBK_CLUSTERS_PER_GROUP");                     // tuning override (never beyond what is resident)
              if (e7 && atoi(e7) > 0) {
                  if (b->clustersPerGroup > 1 && atoi(e7) == 1) { b->clusterSize = 8; resident = ctreeMaxActiveClusters(8); }
                  if (atoi(e7)*groups <= resident) b->clustersPerGroup = atoi(e7);
              }
              if (getenv("SBK_VERBOSE")) std::fprintf(stderr, "sbk plan 5: %d groups, cluster of %d CTAs, %d clusters resident at once, %d per group\n", groups, b->clusterSize, resident, b->clustersPerGroup); }
            const int nwarps = b->clusterSize*8*b->clustersPerGroup;
            // the levels above the cut run on ALL warps of the cluster with barrier.cluster between levels (measured +19% over running
            // them on the first CTA's eight warps with __syncthreads: 6 rounds instead of 10 on the C5 tree)
            int topWarps = nwarps/b->clustersPerGroup, cutWarps = nwarps;
            { const char* e5 = getenv("SBK_TOPWARPS"); if (e5 && atoi(e5) > 0) topWarps = atoi(e5); }      // tuning overrides
            { const char* e6 = getenv("SBK_CUTWARPS"); if (e6 && atoi(e6) > 0) cutWarps = atoi(e6); }
            const sbk::TreeCut cut = sbk::cutTreeForWarps(*t, nwarps, topWarps, cutWarps, b->clustersPerGroup);
            const size_t soBytes = pad16(cut.lists.size()*sizeof(int)), ssBytes = pad16(cut.listStart.size()*sizeof(int));
            std::vector<unsigned char> lv(lblob.size() + soBytes + ssBytes, 0);
            std::memcpy(lv.data(), lblob.data(), lblob.size());
            std::memcpy(lv.data(), cut.bodies.data(), cut.bodies.size()*sizeof(LBody));
            std::memcpy(lv.data() + lblob.size(), cut.lists.data(), cut.lists.size()*sizeof(int));
            std::memcpy(lv.data() + lblob.size() + soBytes, cut.listStart.data(), cut.listStart.size()*sizeof(int));
            a.llistsOff = (uint32_t)lblob.size(); a.llistStartOff = (uint32_t)(lblob.size() + soBytes);
            a.nsub = (int)cut.subStart.size() - 1; a.cutLevel = cut.cutLevel;
            CUDA_TRY(cudaMalloc(&b->dLTablesLevel, lv.size()));
            CUDA_TRY(cudaMemcpyAsync(b->dLTablesLevel, lv.data(), lv.size(), cudaMemcpyHostToDevice, b->stream));
            CUDA_TRY(cudaStreamSynchronize(b->stream));
            a.ltablesLevel = b->dLTablesLevel;
            if (a.treeScratch) cudaFree(a.treeScratch);
            a.treeScratch = nullptr;
            CUDA_TRY(cudaMalloc(&a.treeScratch, ctreeScratchDoubles(n, b->clusterSize*b->clustersPerGroup)*sizeof(double)));
        }
      } }
    if (plan == 5 && !a.ltablesLevel) return fail(SBK_ERR_ARG, "plan 5 needs a model made of Pin / Slider / Universal / Ball / Free mobilizers (quaternion mode)");
    CUDA_TRY(cudaMalloc(&a.cache, (size_t)cacheDoubles*sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(a.cache, 0, (size_t)cacheDoubles*sizeof(double), b->stream));
    CUDA_TRY(launchInitGround(a, b->stream)); b->launches++;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->plan = plan; b->stage = ST_EMPTY; b->abiValid = false; b->accelValid = false;
    return SBK_OK;
}
sbk_batch* sbk_batch_create(const sbk_topology* t, int n, int device, void* stream) {
    if (!t || n < 1) { fail(SBK_ERR_ARG, "sbk_batch_create: bad arguments"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        fail(SBK_ERR_CUDA, "sbk_batch_create: no usable CUDA device " + std::to_string(device) +
                           " (this library has no CPU fallback)");
        return nullptr;
    }
    { cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
          cudaGetLastError();
          fail(SBK_ERR_CUDA, "sbk_batch_create: device " + std::to_string(device) + " is not an sm_100 GPU (the kernels are built for sm_100a only)");
          return nullptr;
      } }
    sbk_batch* b = new sbk_batch();
    b->topo = t; b->N = n; b->device = device;
    auto bail = [&](const std::string& m) { fail(SBK_ERR_CUDA, m); sbk_batch_destroy(b); return (sbk_batch*)nullptr; };
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice failed");
    if (stream) b->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed"); b->ownStream = true; }
    cudaEventCreate(&b->ev0); cudaEventCreate(&b->ev1);

    KArgs& a = b->a; std::memset(&a, 0, sizeof a);
    a.nb = t->nb; a.nq = t->nq; a.nu = t->nu; a.nquat = t->nquat;
    a.gx = t->grav[0]; a.gy = t->grav[1]; a.gz = t->grav[2];
    a.N = n;
    const size_t ny = (size_t)t->nq + t->nu, N = (size_t)n;
    bool ok = true;
    auto dalloc = [&](double** p, size_t doubles) { if (ok && cudaMalloc(p, std::max<size_t>(doubles, 1)*sizeof(double)) != cudaSuccess) ok = false;
                                                    if (ok) cudaMemsetAsync(*p, 0, std::max<size_t>(doubles, 1)*sizeof(double), b->stream); };
    dalloc(&a.y, ny*N); dalloc(&a.ydot, ny*N); dalloc(&a.qdotdot, (size_t)t->nq*N); dalloc(&a.qerr, (size_t)std::max(t->nquat, 1)*N);
    const size_t Nb = (N + BLK_LANES - 1)/BLK_LANES*BLK_LANES;        // integrator vectors are CTA-blocked: [block][slot][lane]
    dalloc(&a.yb, ny*Nb);
    dalloc(&a.y0, ny*Nb); dalloc(&a.f0, ny*Nb); dalloc(&a.fa, ny*Nb); dalloc(&a.fb, ny*Nb); dalloc(&a.ys, ny*Nb);
    dalloc(&a.tcur, N); dalloc(&a.errNorm, N);
    if (!t->twoPoint.empty()) dalloc(&a.f2, (size_t)t->nb*6*N);
    dalloc(&b->dOpA, (size_t)t->nu*N); dalloc(&b->dOpB, (size_t)t->nu*N); dalloc(&b->dOpOut, (size_t)t->nu*N); dalloc(&b->dOpF, (size_t)t->nb*6*N);
    if (ok && cudaMalloc(&a.status, N*sizeof(int)) != cudaSuccess) ok = false;
    { const size_t nblk = (N + BLK_LANES - 1)/BLK_LANES;
      if (ok && cudaMalloc(&a.taskCounter, (2 + nblk)*sizeof(int)) != cudaSuccess) ok = false;   // 64-bit task counter, then blockDone[nblk]
      if (ok) a.blockDone = a.taskCounter + 2; }
    if (ok && cudaMalloc(&a.projCount, N*sizeof(int)) != cudaSuccess) ok = false;
    if (ok && cudaMalloc(&a.lflags, N*sizeof(int)) != cudaSuccess) ok = false;
    if (!ok) return bail(std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    cudaMemsetAsync(a.status, 0, N*sizeof(int), b->stream);
    cudaMemsetAsync(a.projCount, 0, N*sizeof(int), b->stream);
    cudaMemsetAsync(a.lflags, 0, N*sizeof(int), b->stream);
    // default state: q = 0 except quaternions (1,0,0,0), like the reference's default State
    if (t->nquat) {
        std::vector<double> q0((size_t)t->nq*N, 0.0);
        for (int i = 0; i < t->nb; ++i) if (t->quatIndex[i] >= 0) for (size_t k = 0; k < N; ++k) q0[(size_t)t->q0[i]*N + k] = 1.0;
        cudaMemcpyAsync(a.y, q0.data(), q0.size()*sizeof(double), cudaMemcpyHostToDevice, b->stream);
        cudaStreamSynchronize(b->stream);
    }
    sbk_rkm_opts o; sbk_rkm_default_opts(&o);
    a.accuracy = o.accuracy; a.consTol = o.constraint_tol;
    if (configurePlan(b, autoPlan(b)) != SBK_OK) { const std::string m = g_lastError; sbk_batch_destroy(b); g_lastError = m; return nullptr; }
    return b;
}
void sbk_batch_destroy(sbk_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    KArgs& a = b->a;
    void* ptrs[] = {b->dTables, b->dLTables, b->dLTablesLevel, a.treeScratch, a.cache, a.y, a.yb, a.ydot, a.qdotdot, a.qerr, a.y0, a.f0, a.fa, a.fb, a.ys, a.tcur, a.errNorm,
                    b->dOpA, b->dOpB, b->dOpOut, b->dOpF, a.status, a.projCount, b->dScratch,
                    a.hcur, a.lastStep, a.stepsTaken, a.attempts, a.taskCounter, a.lflags, a.f2};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (b->ev0) cudaEventDestroy(b->ev0); if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->ownStream && b->stream) cudaStreamDestroy(b->stream);
    delete b;
}
int sbk_batch_size(const sbk_batch* b) { return b ? b->N : 0; }
int sbk_batch_set_plan(sbk_batch* b, int plan) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (plan == 0) plan = autoPlan(b);
    if (plan < 1 || plan > 5) return fail(SBK_ERR_ARG, "sbk_batch_set_plan: plan must be 0..5");
    if (plan == 5 && !b->topo->localOk) return fail(SBK_ERR_ARG, "sbk_batch_set_plan: plan 5 needs a model made of Pin / Slider / Universal / Ball / Free mobilizers (quaternion mode)");
    if (plan == 2 && !fusedOk(b))
        return fail(SBK_ERR_ARG, "sbk_batch_set_plan: the register-resident fused plan needs a serial chain of 1-2 Pin/Slider mobilizers");
    if (plan == b->plan) return SBK_OK;
    return configurePlan(b, plan);
}
int sbk_batch_get_plan(const sbk_batch* b) { return b ? b->plan : 0; }
int sbk_synchronize(sbk_batch* b) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}

// ---- state ----------------------------------------------------------------------------------
static void invalidate(sbk_batch* b) { b->stage = ST_EMPTY; b->abiValid = false; b->accelValid = false; }
static void stateWasSet(sbk_batch* b) {
    invalidate(b); b->adaptiveInit = false; b->integInit = false;
    // a new state starts a new history: the per-instance status words describe the operations since the last set_state
    cudaMemsetAsync(b->a.status, 0, (size_t)b->N*sizeof(int), b->stream);
}
// Integrator::initialize (Integrator.cpp:367-377): realizeAndProjectKinematicsWithThrow(ForceProjection) normalises every
// quaternion before the first step and counts one q projection (IntegratorRep.h:799-801) when the model has any.
static int integratorInitialize(sbk_batch* b) {
    if (b->integInit) return SBK_OK;
    if (b->topo->nquat > 0) { CUDA_TRY(launchInitProject(b->a, b->stream)); b->launches++; }
    b->integInit = true;
    return SBK_OK;
}

int sbk_set_state(sbk_batch* b, const double* q, const double* u, const double* t) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo;
    if (q) if (int rc = h2d(b, b->a.y, q, T->nq)) return rc;
    if (u) if (int rc = h2d(b, b->a.y + (size_t)T->nq*b->N, u, T->nu)) return rc;
    if (t) CUDA_TRY(cudaMemcpyAsync(b->a.tcur, t, (size_t)b->N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));   // the host buffers may be reused by the caller
    stateWasSet(b);
    return SBK_OK;
}
// Asynchronous variants for PINNED host buffers: the copies are queued on the batch's stream and the call returns at once;
// the caller must not touch the buffers before sbk_synchronize() (or a later synchronous call).  One synchronisation per
// set -> step -> get round trip instead of three (several processes sharing one host serialise on those).
int sbk_set_state_async(sbk_batch* b, const double* q, const double* u, const double* t) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo;
    if (q) if (int rc = h2d(b, b->a.y, q, T->nq)) return rc;
    if (u) if (int rc = h2d(b, b->a.y + (size_t)T->nq*b->N, u, T->nu)) return rc;
    if (t) CUDA_TRY(cudaMemcpyAsync(b->a.tcur, t, (size_t)b->N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
    stateWasSet(b);
    return SBK_OK;
}
int sbk_get_state_async(sbk_batch* b, double* q, double* u, double* t) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo;
    if (q) CUDA_TRY(cudaMemcpyAsync(q, b->a.y, (size_t)T->nq*b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (u) CUDA_TRY(cudaMemcpyAsync(u, b->a.y + (size_t)T->nq*b->N, (size_t)T->nu*b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (t) CUDA_TRY(cudaMemcpyAsync(t, b->a.tcur, (size_t)b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    return SBK_OK;
}
int sbk_get_state(sbk_batch* b, double* q, double* u, double* t) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo;
    if (q) CUDA_TRY(cudaMemcpyAsync(q, b->a.y, (size_t)T->nq*b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (u) CUDA_TRY(cudaMemcpyAsync(u, b->a.y + (size_t)T->nq*b->N, (size_t)T->nu*b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (t) CUDA_TRY(cudaMemcpyAsync(t, b->a.tcur, (size_t)b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}
int sbk_set_state_aos(sbk_batch* b, const double* q, const double* u) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo; const int N = b->N;
    if (int rc = ensureScratch(b, (size_t)std::max(T->nq, T->nu)*N)) return rc;
    if (q) {
        CUDA_TRY(cudaMemcpyAsync(b->dScratch, q, (size_t)T->nq*N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(launchTranspose(b->dScratch, b->a.y, N, T->nq, b->stream)); b->launches++;
    }
    if (u) {
        CUDA_TRY(cudaMemcpyAsync(b->dScratch, u, (size_t)T->nu*N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(launchTranspose(b->dScratch, b->a.y + (size_t)T->nq*N, N, T->nu, b->stream)); b->launches++;
    }
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    stateWasSet(b);
    return SBK_OK;
}
int sbk_get_state_aos(sbk_batch* b, double* q, double* u) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    const sbk_topology* T = b->topo; const int N = b->N;
    if (int rc = ensureScratch(b, (size_t)std::max(T->nq, T->nu)*N)) return rc;
    if (q) {
        CUDA_TRY(launchTranspose(b->a.y, b->dScratch, T->nq, N, b->stream)); b->launches++;
        CUDA_TRY(cudaMemcpyAsync(q, b->dScratch, (size_t)T->nq*N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    if (u) {
        CUDA_TRY(launchTranspose(b->a.y + (size_t)T->nq*N, b->dScratch, T->nu, N, b->stream)); b->launches++;
        CUDA_TRY(cudaMemcpyAsync(u, b->dScratch, (size_t)T->nu*N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return SBK_OK;
}
int sbk_state_device_ptrs(sbk_batch* b, double** q, double** u, double** t) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (q) *q = b->a.y; if (u) *u = b->a.y + (size_t)b->topo->nq*b->N; if (t) *t = b->a.tcur;
    return SBK_OK;
}
int sbk_state_touched(sbk_batch* b) { if (!b) return fail(SBK_ERR_ARG, "null batch"); stateWasSet(b); return SBK_OK; }

// ---- realize ----------------------------------------------------------------------------------
static int realizeKin(sbk_batch* b, int toStage) {
    if (int rc = useDevice(b)) return rc;
    if (b->stage < ST_POSITION) {
        if (int rc = launch(b, OP_KIN)) return rc;
        b->abiValid = false; b->accelValid = false;
    }
    if (b->stage < toStage) b->stage = toStage;
    return SBK_OK;
}
int sbk_realize_position(sbk_batch* b) { if (!b) return fail(SBK_ERR_ARG, "null batch"); return realizeKin(b, ST_POSITION); }
int sbk_realize_velocity(sbk_batch* b) { if (!b) return fail(SBK_ERR_ARG, "null batch"); return realizeKin(b, ST_VELOCITY); }
int sbk_realize_articulated_body_inertias(sbk_batch* b) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = realizeKin(b, ST_POSITION)) return rc;
    if (!b->abiValid) { if (int rc = launch(b, OP_ABI)) return rc; b->abiValid = true; }
    return SBK_OK;
}
int sbk_realize_acceleration(sbk_batch* b) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (b->stage == ST_ACCELERATION && b->accelValid) return SBK_OK;
    // one fused launch: kinematics + forces + ABI + both acceleration sweeps
    b->a.fmobOut = b->dOpA; b->a.FbodyOut = b->dOpF;
    CUDA_TRY(cudaMemsetAsync(b->dOpF, 0, (size_t)6*b->N*sizeof(double), b->stream));   // Ground row
    int rc = launch(b, OP_EVAL);
    b->a.fmobOut = nullptr; b->a.FbodyOut = nullptr;
    if (rc) return rc;
    if (b->a.ntp)      // forces a two-point element applies to Ground are part of getRigidBodyForces()[0]
        CUDA_TRY(cudaMemcpyAsync(b->dOpF, b->a.f2, (size_t)6*b->N*sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
    b->stage = ST_ACCELERATION; b->abiValid = true; b->accelValid = true; b->realizations++;
    return SBK_OK;
}

// ---- getters ----------------------------------------------------------------------------------
int sbk_get_udot(sbk_batch* b, double* udot) {
    if (!b || !udot) return fail(SBK_ERR_ARG, "sbk_get_udot: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_ACCELERATION, "sbk_get_udot")) return rc;
    return d2h(b, udot, b->a.ydot + (size_t)b->topo->nq*b->N, b->topo->nu);
}
int sbk_get_qdot(sbk_batch* b, double* qdot) {
    if (!b || !qdot) return fail(SBK_ERR_ARG, "sbk_get_qdot: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_VELOCITY, "sbk_get_qdot")) return rc;
    return d2h(b, qdot, b->a.ydot, b->topo->nq);
}
int sbk_get_qdotdot(sbk_batch* b, double* qdd) {
    if (!b || !qdd) return fail(SBK_ERR_ARG, "sbk_get_qdotdot: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_ACCELERATION, "sbk_get_qdotdot")) return rc;
    return d2h(b, qdd, b->a.qdotdot, b->topo->nq);
}
int sbk_get_qerr(sbk_batch* b, double* qerr) {
    if (!b || !qerr) return fail(SBK_ERR_ARG, "sbk_get_qerr: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_get_qerr")) return rc;
    if (b->topo->nquat == 0) return SBK_OK;
    return d2h(b, qerr, b->a.qerr, b->topo->nquat);
}
static int getBodyField(sbk_batch* b, int field, int width, double* out, int st, const char* who) {
    if (!b || !out) return fail(SBK_ERR_ARG, std::string(who) + ": null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, st, who)) return rc;
    const size_t rows = (size_t)b->topo->nb*width;
    if (int rc = ensureScratch(b, rows*b->N)) return rc;
    CUDA_TRY(launchGatherBodyField(b->a, field, width, b->dScratch, b->stream)); b->launches++;
    return d2h(b, out, b->dScratch, rows);
}
int sbk_get_body_transforms(sbk_batch* b, double* X)  { return getBodyField(b, F_XGB, 12, X, ST_POSITION, "sbk_get_body_transforms"); }
int sbk_get_body_velocities(sbk_batch* b, double* V)  { return getBodyField(b, F_VGB, 6, V, ST_VELOCITY, "sbk_get_body_velocities"); }
int sbk_get_body_accelerations(sbk_batch* b, double* A) {
    if (b && !b->accelValid) return fail(SBK_ERR_STAGE, "sbk_get_body_accelerations: accelerations are not realized");
    return getBodyField(b, F_AGB, 6, A, ST_ACCELERATION, "sbk_get_body_accelerations");
}
int sbk_get_applied_forces(sbk_batch* b, double* fmob, double* Fbody) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_ACCELERATION, "sbk_get_applied_forces")) return rc;
    if (!b->accelValid) return fail(SBK_ERR_STAGE, "sbk_get_applied_forces: forces were overwritten by an operator call; realize again");
    if (fmob)  if (int rc = d2h(b, fmob, b->dOpA, b->topo->nu)) return rc;
    if (Fbody) if (int rc = d2h(b, Fbody, b->dOpF, (size_t)b->topo->nb*6)) return rc;
    return SBK_OK;
}

int sbk_calc_energy(sbk_batch* b, double* ke, double* pe) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ke ? ST_VELOCITY : ST_POSITION, "sbk_calc_energy")) return rc;
    if (int rc = ensureScratch(b, (size_t)2*b->N)) return rc;
    CUDA_TRY(launchEnergy(b->a, b->dScratch, b->dScratch + b->N, b->stream)); b->launches++;
    if (ke) CUDA_TRY(cudaMemcpyAsync(ke, b->dScratch, (size_t)b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (pe) CUDA_TRY(cudaMemcpyAsync(pe, b->dScratch + b->N, (size_t)b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}

int sbk_calc_mobilizer_reaction_forces(sbk_batch* b, double* FM_G) {
    if (!b || !FM_G) return fail(SBK_ERR_ARG, "sbk_calc_mobilizer_reaction_forces: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_ACCELERATION, "sbk_calc_mobilizer_reaction_forces")) return rc;
    if (!b->accelValid) return fail(SBK_ERR_STAGE, "sbk_calc_mobilizer_reaction_forces: an operator call overwrote the acceleration cache; realize acceleration again");
    const size_t rows = (size_t)b->topo->nb*6;
    if (int rc = ensureScratch(b, rows*b->N)) return rc;
    CUDA_TRY(launchReaction(b->a, b->dScratch, b->stream)); b->launches++;
    return d2h(b, FM_G, b->dScratch, rows);
}
int sbk_multiply_by_system_jacobian(sbk_batch* b, const double* v, double* Jv) {
    if (!b || !v || !Jv) return fail(SBK_ERR_ARG, "sbk_multiply_by_system_jacobian: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_multiply_by_system_jacobian")) return rc;
    const size_t rows = (size_t)b->topo->nb*6;
    if (int rc = ensureScratch(b, rows*b->N)) return rc;
    if (int rc = h2d(b, b->dOpB, v, b->topo->nu)) return rc;
    CUDA_TRY(launchJacobian(b->a, b->dOpB, b->dScratch, b->stream)); b->launches++;
    return d2h(b, Jv, b->dScratch, rows);
}
int sbk_multiply_by_system_jacobian_transpose(sbk_batch* b, const double* F, double* JtF) {
    if (!b || !F || !JtF) return fail(SBK_ERR_ARG, "sbk_multiply_by_system_jacobian_transpose: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_multiply_by_system_jacobian_transpose")) return rc;
    const size_t rows = (size_t)b->topo->nb*6;
    if (int rc = ensureScratch(b, 2*rows*b->N)) return rc;
    if (int rc = h2d(b, b->dScratch, F, rows)) return rc;
    CUDA_TRY(launchJacobianTranspose(b->a, b->dScratch, b->dScratch + rows*b->N, b->dOpOut, b->stream)); b->launches++;
    return d2h(b, JtF, b->dOpOut, b->topo->nu);
}

int sbk_calc_composite_body_inertias(sbk_batch* b, double* R) {
    if (!b || !R) return fail(SBK_ERR_ARG, "sbk_calc_composite_body_inertias: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_calc_composite_body_inertias")) return rc;
    const size_t rows = (size_t)b->topo->nb*10;
    if (int rc = ensureScratch(b, rows*b->N)) return rc;
    CUDA_TRY(launchCompositeBodyInertias(b->a, b->dScratch, b->stream)); b->launches++;
    return d2h(b, R, b->dScratch, rows);
}

// ---- operators --------------------------------------------------------------------------------
int sbk_calc_acceleration(sbk_batch* b, const double* fmob, const double* Fbody, double* udot, double* A_GB) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_VELOCITY, "sbk_calc_acceleration")) return rc;
    if (!b->abiValid) { if (int rc = launch(b, OP_ABI)) return rc; b->abiValid = true; }
    b->accelValid = false;
    if (fmob)  if (int rc = h2d(b, b->dOpB, fmob, b->topo->nu)) return rc;
    if (Fbody) if (int rc = h2d(b, b->dOpF, Fbody, (size_t)b->topo->nb*6)) return rc;
    b->a.fmobIn = fmob ? b->dOpB : nullptr; b->a.FbodyIn = Fbody ? b->dOpF : nullptr; b->a.vecOut = b->dOpOut;
    int rc = launch(b, OP_CALCACC);
    b->a.fmobIn = nullptr; b->a.FbodyIn = nullptr; b->a.vecOut = nullptr;
    if (rc) return rc;
    if (udot) if (int rc2 = d2h(b, udot, b->dOpOut, b->topo->nu)) return rc2;
    if (A_GB) {
        const size_t rows = (size_t)b->topo->nb*6;
        if (int rc2 = ensureScratch(b, rows*b->N)) return rc2;
        CUDA_TRY(launchGatherBodyField(b->a, F_AGB, 6, b->dScratch, b->stream)); b->launches++;
        if (int rc2 = d2h(b, A_GB, b->dScratch, rows)) return rc2;
    }
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}
int sbk_multiply_by_M(sbk_batch* b, const double* v, double* Mv) {
    if (!b || !v || !Mv) return fail(SBK_ERR_ARG, "sbk_multiply_by_M: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_multiply_by_M")) return rc;
    b->accelValid = false;
    if (int rc = h2d(b, b->dOpB, v, b->topo->nu)) return rc;
    b->a.vecIn = b->dOpB; b->a.vecOut = b->dOpOut;
    int rc = launch(b, OP_MULM);
    b->a.vecIn = nullptr; b->a.vecOut = nullptr;
    if (rc) return rc;
    return d2h(b, Mv, b->dOpOut, b->topo->nu);
}
int sbk_multiply_by_MInv(sbk_batch* b, const double* v, double* MinvV) {
    if (!b || !v || !MinvV) return fail(SBK_ERR_ARG, "sbk_multiply_by_MInv: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_POSITION, "sbk_multiply_by_MInv")) return rc;
    if (!b->abiValid) { if (int rc = launch(b, OP_ABI)) return rc; b->abiValid = true; }   // lazy, like the reference
    b->accelValid = false;
    if (int rc = h2d(b, b->dOpB, v, b->topo->nu)) return rc;
    b->a.fmobIn = b->dOpB; b->a.FbodyIn = nullptr; b->a.vecOut = b->dOpOut;   // M^-1: f = v, no body forces
    int rc = launch(b, OP_MULMINV);
    b->a.fmobIn = nullptr; b->a.vecOut = nullptr;
    if (rc) return rc;
    return d2h(b, MinvV, b->dOpOut, b->topo->nu);
}
int sbk_calc_residual_force(sbk_batch* b, const double* fmob, const double* Fbody, const double* knownUdot, double* residual) {
    if (!b || !residual) return fail(SBK_ERR_ARG, "sbk_calc_residual_force: null argument");
    if (int rc = useDevice(b)) return rc;
    if (int rc = needStage(b, ST_VELOCITY, "sbk_calc_residual_force")) return rc;
    b->accelValid = false;
    if (fmob)      if (int rc = h2d(b, b->dOpA, fmob, b->topo->nu)) return rc;
    if (knownUdot) if (int rc = h2d(b, b->dOpB, knownUdot, b->topo->nu)) return rc;
    if (Fbody)     if (int rc = h2d(b, b->dOpF, Fbody, (size_t)b->topo->nb*6)) return rc;
    b->a.fmobIn = fmob ? b->dOpA : nullptr; b->a.FbodyIn = Fbody ? b->dOpF : nullptr;
    b->a.vecIn = knownUdot ? b->dOpB : nullptr; b->a.vecOut = b->dOpOut;
    int rc = launch(b, OP_RESID);
    b->a.fmobIn = nullptr; b->a.FbodyIn = nullptr; b->a.vecIn = nullptr; b->a.vecOut = nullptr;
    if (rc) return rc;
    return d2h(b, residual, b->dOpOut, b->topo->nu);
}

// ---- integrator -------------------------------------------------------------------------------
void sbk_rkm_default_opts(sbk_rkm_opts* o) {
    if (!o) return;
    o->accuracy = 1e-3; o->constraint_tol = 1e-4; o->use_infinity_norm = 0; o->project_every_step = 0;
}
int sbk_rkm_step(sbk_batch* b, double h, int nsteps, const sbk_rkm_opts* opts, double* errNorm) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (!(h > 0) || nsteps < 0) return fail(SBK_ERR_ARG, "sbk_rkm_step: need h > 0 and nsteps >= 0");
    if (int rc = useDevice(b)) return rc;
    sbk_rkm_opts o; sbk_rkm_default_opts(&o);
    if (opts) { o = *opts; if (!(o.accuracy > 0)) o.accuracy = 1e-3; if (!(o.constraint_tol > 0)) o.constraint_tol = o.accuracy/10; }
    KArgs& a = b->a;
    a.h = h; a.nsteps = nsteps; a.accuracy = o.accuracy; a.consTol = o.constraint_tol;
    a.useInfNorm = o.use_infinity_norm; a.projectEveryStep = o.project_every_step;
    if (nsteps > 0) {
        if (int rc = integratorInitialize(b)) return rc;
        CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
        if (b->plan == 2) {
            std::vector<int> joints(b->topo->nb);
            for (int i = 0; i < b->topo->nb; ++i) joints[i] = b->topo->bodies[i].joint;
            CUDA_TRY(launchFusedRkm(a, joints.data(), false, b->stream)); b->launches++;
        } else if (b->plan == 4) {
            CUDA_TRY(launchGlRkm(a, b->stream)); b->launches++;
        } else if (b->plan == 5) {
            CUDA_TRY(launchCtreeRkm(a, b->clusterSize, b->clustersPerGroup, b->stream)); b->launches++;
        } else if (int rc = launch(b, OP_RKM)) return rc;
        CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
        invalidate(b);
        b->stepsTaken += (int64_t)nsteps*b->N; b->realizations += (int64_t)5*nsteps*b->N;
    }
    if (errNorm) {
        CUDA_TRY(cudaMemcpyAsync(errNorm, a.errNorm, (size_t)b->N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return SBK_OK;
}
void sbk_adaptive_default_opts(sbk_adaptive_opts* o) {
    if (!o) return;
    o->accuracy = 1e-3; o->constraint_tol = 1e-4; o->init_step = 0.01; o->min_step = -1; o->max_step = -1;
    o->use_infinity_norm = 0; o->project_every_step = 0; o->allow_interpolation = 0; o->max_attempts = 0;
}
int sbk_rkm_adaptive(sbk_batch* b, double tFinal, const sbk_adaptive_opts* opts, int32_t* steps, int32_t* attempts, double* lastStep) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (b->plan == 3) {
        // the CTA-per-instance plan keeps its records instance-major: error-controlled stepping (one step-size history per
        // instance) borrows the thread-per-instance plan's layout and kernel for the call; the state arrays are shared
        if (int rc = configurePlan(b, 1)) return rc;
        const int rc = sbk_rkm_adaptive(b, tFinal, opts, steps, attempts, lastStep);
        const int rc2 = configurePlan(b, 3);
        return rc ? rc : rc2;
    }
    if (!b->topo->twoPoint.empty()) return fail(SBK_ERR_ARG, "sbk_rkm_adaptive: not available for models with two-point force elements (fixed steps only)");
    sbk_adaptive_opts o; sbk_adaptive_default_opts(&o);
    if (opts) {
        o = *opts;
        if (!(o.accuracy > 0)) o.accuracy = 1e-3;
        if (!(o.constraint_tol > 0)) o.constraint_tol = o.accuracy/10;
        if (!(o.init_step > 0)) o.init_step = 0.01;
    }
    KArgs& a = b->a; const size_t N = (size_t)b->N;
    if (!a.hcur) {
        CUDA_TRY(cudaMalloc(&a.hcur, N*sizeof(double))); CUDA_TRY(cudaMalloc(&a.lastStep, N*sizeof(double)));
        CUDA_TRY(cudaMalloc(&a.stepsTaken, N*sizeof(int))); CUDA_TRY(cudaMalloc(&a.attempts, N*sizeof(int)));
    }
    if (!b->adaptiveInit) {       // Integrator::initialize: step size = initial step, counters cleared
        double h0 = o.init_step;
        if (o.min_step > 0) h0 = std::max(h0, o.min_step);
        if (o.max_step > 0) h0 = std::min(h0, o.max_step);
        std::vector<double> hv(N, h0);
        CUDA_TRY(cudaMemcpyAsync(a.hcur, hv.data(), N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(a.lastStep, hv.data(), N*sizeof(double), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemsetAsync(a.stepsTaken, 0, N*sizeof(int), b->stream));
        CUDA_TRY(cudaMemsetAsync(a.attempts, 0, N*sizeof(int), b->stream));
        b->adaptStepsSeen = 0; b->adaptAttemptsSeen = 0;
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        b->adaptiveInit = true;
    }
    a.tFinal = tFinal; a.accuracy = o.accuracy; a.consTol = o.constraint_tol; a.minStep = o.min_step; a.maxStep = o.max_step;
    a.useInfNorm = o.use_infinity_norm; a.projectEveryStep = o.project_every_step; a.allowInterp = o.allow_interpolation;
    a.maxAttempts = o.max_attempts > 0 ? o.max_attempts : 1000000;
    if (int rc = integratorInitialize(b)) return rc;
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    if (b->plan == 2) {
        std::vector<int> joints(b->topo->nb);
        for (int i = 0; i < b->topo->nb; ++i) joints[i] = b->topo->bodies[i].joint;
        CUDA_TRY(launchFusedRkm(a, joints.data(), true, b->stream)); b->launches++;
    } else {
        // plans 1 and 4 share the record layout: error-controlled stepping always runs the thread-per-instance kernel
        // (every instance keeps its own step-size history; the grid-level plan steps the whole batch in lockstep)
        CUDA_TRY(launchTpi(OP_RKM_ADAPT, a, b->stream)); b->launches++;
    }
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    invalidate(b);
    {   // Integrator::getNumStepsTaken / getNumRealizations over the batch: one fresh evaluation per step + 4 per attempt
        std::vector<int> hs(N), ha(N);
        CUDA_TRY(cudaMemcpyAsync(hs.data(), a.stepsTaken, N*sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaMemcpyAsync(ha.data(), a.attempts, N*sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        int64_t ss = 0, sa = 0; for (size_t k = 0; k < N; ++k) { ss += hs[k]; sa += ha[k]; }
        b->stepsTaken += ss - b->adaptStepsSeen; b->realizations += (ss - b->adaptStepsSeen) + 4*(sa - b->adaptAttemptsSeen);
        b->adaptStepsSeen = ss; b->adaptAttemptsSeen = sa;
        if (steps)    std::memcpy(steps, hs.data(), N*sizeof(int));
        if (attempts) std::memcpy(attempts, ha.data(), N*sizeof(int));
    }
    if (lastStep) CUDA_TRY(cudaMemcpyAsync(lastStep, a.lastStep, N*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return SBK_OK;
}
int sbk_rkm_stats(sbk_batch* b, int64_t* steps, int64_t* realizations, int64_t* qproj) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    if (steps) *steps = b->stepsTaken; if (realizations) *realizations = b->realizations;
    if (qproj) {
        std::vector<int> h(b->N);
        CUDA_TRY(cudaMemcpyAsync(h.data(), b->a.projCount, (size_t)b->N*sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        int64_t s = 0; for (int v : h) s += v; *qproj = s;
    }
    return SBK_OK;
}
int sbk_get_status(sbk_batch* b, int32_t* status, int64_t* nbad) {
    if (!b) return fail(SBK_ERR_ARG, "null batch");
    if (int rc = useDevice(b)) return rc;
    std::vector<int> h(b->N);
    CUDA_TRY(cudaMemcpyAsync(h.data(), b->a.status, (size_t)b->N*sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    int64_t bad = 0; for (int k = 0; k < b->N; ++k) { if (h[k]) ++bad; if (status) status[k] = h[k]; }
    if (nbad) *nbad = bad;
    return SBK_OK;
}
int64_t sbk_launch_count(const sbk_batch* b) { return b ? b->launches : 0; }
// Name of the fixed-step integrator kernel the batch's plan launches, spelled like the demangled name a profiler prints
// (bench.py matches its live runs with the committed ncu captures by it).
int sbk_integrator_kernel_name(const sbk_batch* b, char* buf, int cap) {
    if (!b || !buf || cap < 1) return fail(SBK_ERR_ARG, "sbk_integrator_kernel_name: bad arguments");
    std::string s;
    if (b->plan == 2) s = "fusedRkmKernel";
    else if (b->plan == 3) s = "lpKernel<7>";
    else if (b->plan == 4) s = "glRkmKernel";
    else if (b->plan == 5) s = "ctreeRkmKernel";
    else {
        const KArgs& a = b->a; const int m = a.jointMask;
        int jm, minb = 2, stage, vc = 1;
        if (a.ltables) {
            jm = (((m & ~JM_PIN) == 0) ? JM_PIN : JM_MOBILE5) | JM_LOCAL; stage = a.lstageInSmem ? 1 : 0;
            // (same rule as launchTpi) Pin-only models: three work groups per CTA over one staged copy of the tables
            if ((m & ~JM_PIN) == 0 && stage && a.ltableBytes + launchTpiRkmLocalPin_m3_workBytes() + 2048 <= (size_t)227*1024) { minb = 1; vc = 3; }
        }
        else { jm = JM_ALL; stage = a.stageInSmem ? 1 : 0; }
        s = "tpiKernel<7, " + std::to_string(stage) + ", " + std::to_string(minb) + ", " + std::to_string(jm) + ", " + std::to_string(vc) + ">";
    }
    std::snprintf(buf, (size_t)cap, "%s", s.c_str());
    return SBK_OK;
}
double sbk_last_kernel_ms(const sbk_batch* b) {
    if (!b || !b->ev0 || !b->ev1) return 0;
    cudaSetDevice(b->device);
    if (cudaEventSynchronize(b->ev1) != cudaSuccess) { cudaGetLastError(); return 0; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

// Memory-pattern probe (diagnostics): GB/s delivered for the record access pattern of plan 1.
int sbk_mem_pattern_probe(int device, int n, int nb, int rows_in, int rows_out, int sweeps, int min_blocks, double* ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return fail(SBK_ERR_CUDA, "cudaSetDevice failed");
    if (rows_in > 48 || rows_in < 0 || rows_out < 0) return fail(SBK_ERR_ARG, "rows_in must be 0..48");
    double* d = nullptr; const size_t doubles = (size_t)nb*(rows_in + rows_out)*n;
    CUDA_TRY(cudaMalloc(&d, doubles*sizeof(double)));
    cudaMemset(d, 0, doubles*sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launchMemPattern(d, n, nb, rows_in, rows_out, 1, min_blocks, 0);
    cudaEventRecord(e0, 0);
    cudaError_t le = launchMemPattern(d, n, nb, rows_in, rows_out, sweeps, min_blocks, 0);
    cudaEventRecord(e1, 0);
    cudaError_t se = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (le != cudaSuccess || se != cudaSuccess) return fail(SBK_ERR_CUDA, "mem pattern probe failed");
    if (ms_out) *ms_out = ms;
    return SBK_OK;
}

// FP64 roofline probe (bench only): executes blocks*threads*iters*8 DFMAs on the batch's device.
int sbk_dfma_probe(int device, int blocks, int threads, int iters, double* ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return fail(SBK_ERR_CUDA, "cudaSetDevice failed");
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks*threads*sizeof(double)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launchDfmaProbe(d, 16, blocks, threads, 0);            // warm-up
    cudaEventRecord(e0, 0);
    cudaError_t le = launchDfmaProbe(d, iters, blocks, threads, 0);
    cudaEventRecord(e1, 0);
    cudaError_t se = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (le != cudaSuccess || se != cudaSuccess) return fail(SBK_ERR_CUDA, "dfma probe failed");
    if (ms_out) *ms_out = ms;
    return SBK_OK;
}

} // extern "C"
