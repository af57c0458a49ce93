// sbk_tpi.cuh -- the thread-per-instance kernel template (API operations and the integrator), shared by the
// translation units that instantiate it: sbk_kernels.cu (API operations) and sbk_rkm_{pin,light,mobile5,all}.cu
// (one integrator kernel family per set of mobilizer kinds, compiled in parallel).
#pragma once
#include <algorithm>
#include <type_traits>
#include "sbk_kernels.cuh"
#include "sbk_lrkm.cuh"

namespace sbkd {
namespace {

#ifndef SBK_TPI_THREADS
#define SBK_TPI_THREADS 128
#endif
#ifndef SBK_TPI_MINBLOCKS
#define SBK_TPI_MINBLOCKS 2
#endif
#ifndef SBK_HEAVY_MINB
#define SBK_HEAVY_MINB 2
#endif
constexpr int TPI_THREADS = SBK_TPI_THREADS;
constexpr int SBK_CARRY_STRIDE_DEVICE = 128;

__device__ __forceinline__ uint32_t smemAddr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage `bytes` (multiple of 16) from global to shared with one TMA bulk copy.
__device__ __forceinline__ void tmaStage(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
    const uint32_t bar = smemAddr(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar) : "memory");
    }
}

__device__ __forceinline__ void fillCtx(Ctx& c, const KArgs& a, const unsigned char* tables, bool integrator) {
    c.bodies   = reinterpret_cast<const BodyConst*>(tables);
    c.children = reinterpret_cast<const int*>(tables + a.childrenOff);
    c.forces   = reinterpret_cast<const ForceConst*>(tables + a.forcesOff);
    c.nb = a.nb; c.nq = a.nq; c.nu = a.nu; c.nquat = a.nquat;
    c.gx = a.gx; c.gy = a.gy; c.gz = a.gz;
    c.cache = a.cache; c.cStride = a.cStride; c.cInstStride = a.cInstStride; c.cSpan = a.cSpan; c.cShift = a.cShift; c.cMask = a.cMask;
    c.sStride = a.N; c.sInstStride = 1; c.sSpan = (long long)(a.nq + a.nu)*BLK_LANES;
    c.q = a.y; c.u = a.y + (long long)a.nq*a.N;
    c.qdot = a.ydot; c.udot = a.ydot ? a.ydot + (long long)a.nq*a.N : nullptr;
    c.qdotdot = a.qdotdot; c.qerr = a.qerr;
    c.fmobIn = a.fmobIn; c.FbodyIn = a.FbodyIn; c.fmobOut = a.fmobOut; c.FbodyOut = a.FbodyOut;
    c.vecIn = a.vecIn; c.vecOut = a.vecOut;
    c.status = a.status;
    c.tp = reinterpret_cast<const TwoPointConst*>(tables + a.tpOff); c.ntp = a.ntp; c.f2 = a.f2;
    if (integrator) { c.qdotdot = nullptr; c.qerr = nullptr; c.fmobOut = nullptr; c.FbodyOut = nullptr; }
}
// Thread-per-instance integrator kernels work on a CTA-blocked copy of the state ([block][slot][lane]).
__device__ __forceinline__ void useBlockedState(Ctx& c, const KArgs& a) { c.q = a.yb; c.u = a.yb + (long long)a.nq*BLK_LANES; }
__device__ __forceinline__ void stateToBlocked(const Ctx& c, const KArgs& a, int inst) {
    const int ny = a.nq + a.nu;
#pragma unroll 8
    for (int i = 0; i < ny; ++i) a.yb[stateIndex<true>(c, inst, i)] = __ldcg(a.y + (long long)i*a.N + inst);
}
__device__ __forceinline__ void stateFromBlocked(const Ctx& c, const KArgs& a, int inst, const double* yb = nullptr) {
    const int ny = a.nq + a.nu;
    if (!yb) yb = a.yb;
#pragma unroll 8
    for (int i = 0; i < ny; ++i) __stcg(a.y + (long long)i*a.N + inst, yb[stateIndex<true>(c, inst, i)]);
}
__device__ __forceinline__ LRkmWork fusedWork(const KArgs& a) {
    LRkmWork w; w.Y = a.yb; w.W = a.ys; w.F0 = a.f0; w.F2 = a.fa; w.F3 = a.fb; w.Ynext = a.yb;
    w.accuracy = a.accuracy; w.consTol = a.consTol; w.useInfNorm = a.useInfNorm; w.projectEveryStep = a.projectEveryStep;
    return w;
}

template <class TBL> __device__ __forceinline__ TBL makeTables(const unsigned char* base, const KArgs& a);
template <> __device__ __forceinline__ Tables makeTables<Tables>(const unsigned char* base, const KArgs& a) {
    Tables T; T.bodies = reinterpret_cast<const BodyConst*>(base); T.children = reinterpret_cast<const int*>(base + a.childrenOff);
    T.forces = reinterpret_cast<const ForceConst*>(base + a.forcesOff); return T;
}
template <> __device__ __forceinline__ LTables makeTables<LTables>(const unsigned char* base, const KArgs& a) {
    LTables T; T.bodies = reinterpret_cast<const LBody*>(base); T.children = reinterpret_cast<const int*>(base + a.lchildrenOff);
    T.forces = reinterpret_cast<const ForceConst*>(base + a.lforcesOff); T.fcoef = reinterpret_cast<const double*>(base + a.lfcoefOff); return T;
}

// MINB = resident CTAs per SM the register allocation is sized for: 4 (128 registers) suits
// models made of 1-2 dof mobilizers, 2 (255 registers) models with Ball/Free bodies, whose 3x3 /
// 6x6 articulated-inertia algebra would otherwise spill (measured: +46% on the humanoid).
// VC > 1 (fixed-step integrator only): one CTA of VC*128 threads runs VC independent 128-thread work groups ("virtual CTAs":
// own task slot, own named barrier, own work columns) over ONE staged copy of the tables -- more resident warps per SM than
// VC separate CTAs would fit once every CTA carries its own copy.
template <int OP, bool STAGE, int MINB, int JMASK = JM_ALL, int VC = 1>
__global__ void __launch_bounds__(TPI_THREADS*VC, MINB) tpiKernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ Ctx sctx;                      // ONE context per CTA, read with LDS by every body step
    constexpr bool INTEG = OP == OP_RKM || OP == OP_RKM_ADAPT;
    constexpr bool LOCAL = INTEG && (JMASK & JM_LOCAL) != 0;        // body-frame sweeps with their own (smaller) tables
    typedef std::conditional_t<LOCAL, LTables, Tables> TBL;
    const unsigned char* tables = LOCAL ? a.ltables : a.tables;
    const uint32_t tableBytes = LOCAL ? a.ltableBytes : a.tableBytes;
    if constexpr (STAGE) { tmaStage(smem, tables, tableBytes, &mbar); tables = smem; }
    if (!INTEG && threadIdx.x == 0) fillCtx(sctx, a, tables, false);
    __syncthreads();
    if constexpr (OP == OP_RKM) {
        // Fixed-step integrator: PERSISTENT CTAs (the grid is what fits the machine at once) work through
        // the (block of 128 instances, step) tasks step-major.  A batch whose CTA count is not a multiple of
        // the resident slots (65536 instances = 512 CTAs on 296 slots) then costs nsteps*512/296 rounds
        // instead of nsteps*2.  Steps of one block are ordered through blockDone[block] (release after
        // the step, acquire before the next one); tasks are taken in order, so the task a CTA waits for
        // is always held by a running CTA.  Two ways of taking them: a global counter (work groups of
        // the Pin-only kernel), or -- roundSync, 128-thread CTAs -- round-robin with a grid barrier before
        // every round: the CTAs then stay in the same phase of the step, which is what the memory system
        // likes (humanoid: every launch 49.7 ms; with the counter 49.2 or 53.8 ms depending on how the CTAs
        // happened to drift apart, and a deliberate start offset made every launch slow).  The grid is
        // launched cooperatively, so all its CTAs are resident.
        __shared__ long long sTaskV[VC];     // 64-bit: blocks x steps can exceed 2^31
        const int vc = VC > 1 ? (int)threadIdx.x/TPI_THREADS : 0, tid = VC > 1 ? (int)threadIdx.x - vc*TPI_THREADS : (int)threadIdx.x;
        long long& sTask = sTaskV[vc];
        auto groupSync = [&]() { if constexpr (VC > 1) asm volatile("bar.sync %0, %1;" :: "r"(vc + 1), "n"(TPI_THREADS) : "memory"); else __syncthreads(); };
        Ctx lctx; fillCtx(lctx, a, tables, true); useBlockedState(lctx, a);
        const Ctx& c = lctx;
        const TBL T = makeTables<TBL>(tables, a);
        constexpr int rowsPerGroup = LOCAL ? (int)LFCARRY_ROWS : (int)CARRY_ROWS;
        double* cy = reinterpret_cast<double*>(smem + (STAGE ? tableBytes : 0)) + (size_t)vc*rowsPerGroup*TPI_THREADS + tid;
        RkmWork w;
        w.y = a.yb; w.y0 = a.y0; w.f0 = a.f0; w.fa = a.fa; w.fb = a.fb; w.ys = a.ys;
        w.accuracy = a.accuracy; w.consTol = a.consTol; w.useInfNorm = a.useInfNorm; w.projectEveryStep = a.projectEveryStep;
        const int nblk = (a.N + TPI_THREADS - 1)/TPI_THREADS; const long long total = (long long)nblk*a.nsteps;
        long long myRound = 0;
#pragma unroll 1
        for (;;) {
            if (tid == 0) {
                long long t;
                if (VC == 1 && a.roundSync) {
                    // arrive on the 64-bit counter, wait for the whole grid
                    unsigned long long* bar = reinterpret_cast<unsigned long long*>(a.taskCounter);
                    if (myRound > 0) {
                        __threadfence();
                        atomicAdd(bar, 1ULL);
                        const unsigned long long target = (unsigned long long)myRound*gridDim.x;
                        unsigned long long seen;
                        do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory"); } while (seen < target);
                    }
                    t = myRound*gridDim.x + blockIdx.x; ++myRound;
                    if ((myRound - 1)*(long long)gridDim.x >= total) t = total; else if (t >= total) t = -1;   // -1: idle this round, keep the barrier
                } else
                t = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(a.taskCounter), 1ULL);
                if (t >= 0 && t < total) {
                    const int blk = (int)(t % nblk), step = (int)(t / nblk); int done;
                    do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(a.blockDone + blk) : "memory"); if (done < step) __nanosleep(200); } while (done < step);
                }
                sTask = t;
            }
            groupSync();
            const long long t = sTask;
            if (t >= total) break;
            if (t < 0) continue;
            const int blk = (int)(t % nblk), step = (int)(t / nblk);
            const int inst = blk*TPI_THREADS + tid;
            if (inst < a.N) {
                if (step == 0) stateToBlocked(c, a, inst);
                RkmStepResult r;
                if constexpr (LOCAL) {
                    // fused two-sweep step (sbk_lrkm.cuh).  The velocity buffer alternates with every outward sweep (5 per step);
                    // the data left by the previous step are current unless this is the first step of a launch or that step
                    // projected quaternions (lflags, written by whichever CTA ran it).
                    LRkmState st; st.vb = (step & 1) ? LR_VBUF : 0; st.velValid = step > 0 && a.lflags[inst] != 0;
                    r = lRkmAttempt<JMASK>(c, T, inst, fusedWork(a), a.h, cy, st);
                    a.lflags[inst] = st.velValid ? 1 : 0;
                } else r = tpiRkmStep<true, JMASK>(c, T, inst, w, a.h, cy);
                a.tcur[inst] += a.h;
                if (r.projected) a.projCount[inst] += 1;
                if (step == a.nsteps - 1) {
                    stateFromBlocked(c, a, inst);
                    a.errNorm[inst] = r.errNorm;
                }
                if (a.status && !finiteNorm(r.errNorm)) atomicOr(a.status + inst, 1);   // non-finite error norm
            }
            __threadfence();
            groupSync();
            if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.blockDone + blk), "r"(step + 1) : "memory");
        }
        return;
    }
    const int inst0 = blockIdx.x*blockDim.x + threadIdx.x;
    const bool live = inst0 < a.N;              // the error-controlled kernel votes CTA-wide: its surplus threads stay, predicated off
    if (!live && OP != OP_RKM_ADAPT) return;
    const int inst = live ? inst0 : a.N - 1;
    // API kernels read the context from shared memory; the integrator kernels keep it as a local whose
    // members are kernel parameters (constant bank operands) or one add away from them -- reading
    // it from shared memory cost a generic load plus descriptor moves per access (ncu).
    Ctx lctx;
    if constexpr (INTEG) { fillCtx(lctx, a, tables, true); useBlockedState(lctx, a); }
    const Ctx& c = INTEG ? lctx : sctx;
    const TBL T = makeTables<TBL>(tables, a);  // address space known at compile time: shared if staged, else global
    // LEAN carry: a [CARRY_ROWS][128] block of shared memory behind the staged tables
    double* cy = nullptr;
    if constexpr (OP == OP_RKM || OP == OP_RKM_ADAPT) cy = reinterpret_cast<double*>(smem + (STAGE ? tableBytes : 0)) + threadIdx.x;

    if constexpr (OP == OP_KIN) {
        tpiKinematics<true>(c, inst, c.qdot);
    } else if constexpr (OP == OP_ABI) {
        tpiInward<IN_ABI, true>(c, inst);
    } else if constexpr (OP == OP_EVAL) {
        tpiEvalDerivatives<false, JM_ALL, true>(c, tablesOf(c), inst, cy, c.qdot, c.udot, c.qdotdot);
    } else if constexpr (OP == OP_CALCACC) {
        tpiInward<IN_Z | IN_BIAS, true>(c, inst);
        tpiOutward<true, true>(c, inst, c.vecOut, nullptr);
    } else if constexpr (OP == OP_MULM) {
        for (int b = 1; b < c.nb; ++b) idOutDispatch<false, true>(c, b, inst);
        for (int b = c.nb - 1; b >= 1; --b) idInDispatch<false, true>(c, b, inst);
    } else if constexpr (OP == OP_MULMINV) {
        tpiInward<IN_Z, true>(c, inst);             // c.fmobIn == a.vecIn, c.FbodyIn == null (set by the host)
        tpiOutward<false, true>(c, inst, c.vecOut, nullptr);
    } else if constexpr (OP == OP_RESID) {
        for (int b = 1; b < c.nb; ++b) idOutDispatch<true, true>(c, b, inst);
        for (int b = c.nb - 1; b >= 1; --b) idInDispatch<true, true>(c, b, inst);
    } else if constexpr (OP == OP_RKM_ADAPT) {
        RkmWork w;
        w.y = a.yb; w.y0 = a.y0; w.f0 = a.f0; w.fa = a.fa; w.fb = a.fb; w.ys = a.ys;
        w.accuracy = a.accuracy; w.consTol = a.consTol; w.useInfNorm = a.useInfNorm; w.projectEveryStep = a.projectEveryStep;
        if (live) stateToBlocked(c, a, inst);
        StepLimits lim; lim.accuracy = a.accuracy; lim.minStep = a.minStep; lim.maxStep = a.maxStep;
        AdaptiveState st; st.t = a.tcur[inst]; st.h = a.hcur[inst]; st.lastStep = a.lastStep[inst]; st.steps = 0; st.attempts = 0;
        double lastErr = a.errNorm[inst]; int nproj = 0;
        if constexpr (LOCAL) {
            LRkmWork lw = fusedWork(a); lw.Ynext = a.y0;             // y1 goes to the second state buffer; the two swap on acceptance
            LRkmState ls; ls.vb = 0; ls.velValid = false;
            lRkmAdaptive<JMASK, CtaVote>(c, T, inst, lw, lim, a.tFinal, a.allowInterp, a.maxAttempts, st, cy, ls, lastErr, nproj, live);
            if (live) stateFromBlocked(c, a, inst, lw.Y);
        } else {
            tpiRkmAdaptive<true, JMASK, TBL, CtaVote>(c, T, inst, w, lim, a.tFinal, a.allowInterp, a.maxAttempts, st, cy, lastErr, nproj, live);
            if (live) stateFromBlocked(c, a, inst);
        }
        if (!live) return;
        a.tcur[inst] = st.t; a.hcur[inst] = st.h; a.lastStep[inst] = st.lastStep;
        a.stepsTaken[inst] += st.steps; a.attempts[inst] += st.attempts;
        a.errNorm[inst] = lastErr; a.projCount[inst] += nproj;
        if (a.status && st.t < a.tFinal) atomicOr(a.status + inst, 4);              // attempt budget exhausted
    }
}

// Launch one thread-per-instance operation.  JMASK only matters for the integrator kernels.
template <int OP, int MINB, int JMASK, int VC = 1>
cudaError_t launchOp(const KArgs& a, cudaStream_t stream) {
    static_assert(VC == 1 || OP == OP_RKM, "virtual CTAs exist for the persistent fixed-step kernel only");
    static_assert(TPI_THREADS == SBK_CARRY_STRIDE_DEVICE, "carry columns are laid out for 128-thread CTAs");
    const int grid = (a.N + TPI_THREADS - 1)/TPI_THREADS;
    constexpr bool LOCAL = (OP == OP_RKM || OP == OP_RKM_ADAPT) && (JMASK & JM_LOCAL) != 0;
    const size_t carryBytes = (OP == OP_RKM || OP == OP_RKM_ADAPT) ? (size_t)(LOCAL ? (int)LFCARRY_ROWS : (int)CARRY_ROWS)*TPI_THREADS*sizeof(double) : 0;
    bool stage = LOCAL ? a.lstageInSmem != 0 : a.stageInSmem != 0;
    // builds for more than two resident CTAs stage the tables only if that many copies fit next to the work columns
    if (MINB*VC > 2 && (size_t)MINB*((LOCAL ? a.ltableBytes : a.tableBytes) + VC*carryBytes + 2048) > (size_t)227*1024) stage = false;
    const size_t smemBytes = (stage ? (LOCAL ? a.ltableBytes : a.tableBytes) : 0) + VC*carryBytes;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
        if (e != cudaSuccess) return e;
        int g = grid;
        if constexpr (OP == OP_RKM) {             // persistent: as many CTAs as are resident at once
            int dev = 0, sms = 0, perSm = 0;
            cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, TPI_THREADS*VC, smemBytes);
            if (e != cudaSuccess) return e;
            if (perSm < 1) return cudaErrorLaunchOutOfResources;
            g = std::min((grid + VC - 1)/VC, sms*perSm);
            e = cudaMemsetAsync(a.taskCounter, 0, sizeof(int)*(size_t)(2 + grid), stream);   // 64-bit counter + blockDone[grid]
            if (e != cudaSuccess) return e;
        }
        if constexpr (OP == OP_RKM && VC == 1) {
            KArgs k = a; k.roundSync = g < grid ? 1 : 0;      // every CTA owns its block when the whole batch is resident: nothing to meet for
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)g); cfg.blockDim = dim3(TPI_THREADS); cfg.dynamicSmemBytes = smemBytes; cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
            static const bool noCoop = []() { const char* e = getenv("SBK_NOCOOP"); return e && atoi(e); }();   // diagnostics: plain launch
            cfg.attrs = attr; cfg.numAttrs = (k.roundSync && !noCoop) ? 1 : 0;
            cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, k);
            if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported) {
                // the whole grid cannot be resident right now (shared GPU, MPS limits): fall back to the task counter, which needs no
                // co-residency
                cudaGetLastError();
                k.roundSync = 0; cfg.numAttrs = 0;
                le = cudaLaunchKernelEx(&cfg, kernel, k);
            }
            return le;
        } else {
            kernel<<<g, TPI_THREADS*VC, smemBytes, stream>>>(a);
            return cudaGetLastError();
        }
    };
    return stage ? go(tpiKernel<OP, true, MINB, JMASK, VC>) : go(tpiKernel<OP, false, MINB, JMASK, VC>);
}
// Body of one integrator translation unit (sbk_rkm_<variant>.cu)
// the fixed-step kernel as ONE CTA of VC work groups per SM (there is no error-controlled build of this form)
#define SBK_DEFINE_RKM_VARIANT_VC(NAME, VC, JMASK)                                                  \
    namespace sbkd {                                                                                \
    cudaError_t NAME(KernelOp op, const KArgs& a, cudaStream_t stream) {                            \
        if (op == OP_RKM)       return launchOp<OP_RKM, 1, JMASK, VC>(a, stream);                   \
        return cudaErrorInvalidValue;                                                               \
    }                                                                                               \
    size_t NAME##_workBytes() { return (size_t)VC*LFCARRY_ROWS*TPI_THREADS*sizeof(double); } }
#define SBK_DEFINE_RKM_VARIANT(NAME, MINB, JMASK)                                                   \
    namespace sbkd {                                                                                \
    cudaError_t NAME(KernelOp op, const KArgs& a, cudaStream_t stream) {                            \
        if (op == OP_RKM)       return launchOp<OP_RKM, MINB, JMASK>(a, stream);                    \
        if (op == OP_RKM_ADAPT) return launchOp<OP_RKM_ADAPT, MINB, JMASK>(a, stream);              \
        return cudaErrorInvalidValue;                                                               \
    } }

} // namespace
} // namespace sbkd
