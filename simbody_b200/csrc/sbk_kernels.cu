// sbk_kernels.cu -- sm_100a kernels.
//
// Thread-per-instance plans (1, 2, and the API operations of 4): one thread = one instance; a warp = 32
// instances walking the SAME body at the same time, so the joint-type switch is warp-uniform and every
// access to the CTA-blocked records / state ([block of 128][row][lane]) is a fully coalesced 256-byte warp
// transaction.  Batch-shared body constants are staged into shared memory once per CTA by a TMA bulk copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) and then read as warp broadcasts.
// The fixed-step integrator kernel is persistent (task queue over block x step).  Plan 3 maps a CTA to an
// instance and threads to the bodies of a level; plan 4 maps the whole (cooperative) grid to one tree level
// of the batch.  FP64 throughout; no tensor cores (6x6 spatial operators are not a dense contraction).
#include <algorithm>
#include <math_constants.h>
#include "sbk_kernels.cuh"
#include "sbk_fused.cuh"
#include "sbk_tpi.cuh"

namespace sbkd {

namespace {

// Register-resident fused plan: the whole multi-step RKM loop of one instance in one thread.
template <class E, bool ADAPT>
// three resident CTAs per SM (168 registers, 176 bytes of spill code per thread): 12 warps per SM cover the FP64 latency better than
// 8 at 244 registers -- C2 3.26e9 -> 3.42e9 instance-steps/s; four CTAs at 128 registers: 3.32e9
#ifndef SBK_FUSED_MINB
#define SBK_FUSED_MINB 3
#endif
__global__ void __launch_bounds__(128, SBK_FUSED_MINB) fusedRkmKernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    tmaStage(smem, a.tables, a.tableBytes, &mbar);
    const int inst = blockIdx.x*blockDim.x + threadIdx.x;
    if (inst >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(smem);
    E e; e.b0 = &bodies[1]; e.b1 = &bodies[E::NB - 1];
    e.forces = reinterpret_cast<const ForceConst*>(smem + a.forcesOff);
    e.gx = a.gx; e.gy = a.gy; e.gz = a.gz;
    double y[E::NY];
#pragma unroll
    for (int i = 0; i < E::NY; ++i) y[i] = a.y[(long long)i*a.N + inst];
    double err = 0;
    if constexpr (ADAPT) {
        StepLimits lim; lim.accuracy = a.accuracy; lim.minStep = a.minStep; lim.maxStep = a.maxStep;
        AdaptiveState st; st.t = a.tcur[inst]; st.h = a.hcur[inst]; st.lastStep = a.lastStep[inst]; st.steps = 0; st.attempts = 0;
        err = a.errNorm[inst];
        fusedRkmAdaptive(e, y, lim, a.tFinal, a.allowInterp, a.maxAttempts, a.useInfNorm, st, err);
        a.tcur[inst] = st.t; a.hcur[inst] = st.h; a.lastStep[inst] = st.lastStep;
        a.stepsTaken[inst] += st.steps; a.attempts[inst] += st.attempts;
        if (a.status && st.t < a.tFinal) a.status[inst] |= 4;
    } else {
        for (int s = 0; s < a.nsteps; ++s) err = fusedRkmStep(e, y, a.h, a.useInfNorm);
        a.tcur[inst] += a.nsteps*a.h;
    }
#pragma unroll
    for (int i = 0; i < E::NY; ++i) a.y[(long long)i*a.N + inst] = y[i];
    a.errNorm[inst] = err;
    if (a.status && !finiteNorm(err)) a.status[inst] |= 1;
}
template <class E>
cudaError_t launchFusedT(const KArgs& a, bool adaptive, cudaStream_t stream) {
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.tableBytes);
        if (e != cudaSuccess) return e;
        kernel<<<(a.N + 127)/128, 128, a.tableBytes, stream>>>(a);
        return cudaGetLastError();
    };
    return adaptive ? go(fusedRkmKernel<E, true>) : go(fusedRkmKernel<E, false>);
}

//==============================================================================================
// Level-parallel plan: CTA = instance, threads = bodies of one tree level.
//==============================================================================================
constexpr int LP_THREADS = 256;

__device__ __forceinline__ double blockReduce(double v, bool isMax, double* red) {
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = isMax ? normMax(v, t) : v + t; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = isMax ? 0.0 : 0.0;
    for (int w = 0; w < LP_THREADS/32; ++w) r = isMax ? normMax(r, red[w]) : r + red[w];
    return r;
}

struct LpLevels { const int* order; const int* start; int nlevels; };

template <class F> __device__ __forceinline__ void lpOutward(const LpLevels& L, F f) {
    for (int l = 1; l < L.nlevels; ++l) {
        for (int i = L.start[l] + threadIdx.x; i < L.start[l+1]; i += LP_THREADS) f(L.order[i]);
        __syncthreads();
    }
}
template <class F> __device__ __forceinline__ void lpInward(const LpLevels& L, F f) {
    for (int l = L.nlevels - 1; l >= 1; --l) {
        for (int i = L.start[l] + threadIdx.x; i < L.start[l+1]; i += LP_THREADS) f(L.order[i]);
        __syncthreads();
    }
}
__device__ __forceinline__ void lpEval(const Ctx& c, const int inst, const LpLevels& L, double* qdotDst, double* udotDst, double* qddDst) {
    lpOutward(L, [&](int b) { kinDispatch(c, b, inst, qdotDst); });
    if (c.ntp) { if (threadIdx.x == 0) twoPointPass(c, inst); __syncthreads(); }       // element order, one thread: deterministic sums
    lpInward(L,  [&](int b) { inwardDispatch<IN_ABI | IN_Z | IN_BIAS | IN_FORCES>(c, b, inst); });
    lpOutward(L, [&](int b) { outwardDispatch<true>(c, b, inst, udotDst, qddDst); });
}

// Error norm of IntegratorRep::calcErrorNorm, threads over slots / bodies (cf. rkmErrorNorm).
__device__ double lpErrorNorm(const Ctx& c, const int inst, const KArgs& a, double* red) {
    const int nq = c.nq, nu = c.nu; const bool inf = a.useInfNorm != 0;
    double uAcc = 0, qAcc = 0;
    for (int i = threadIdx.x; i < nu; i += LP_THREADS) {
        const double u0 = fabs(ldS<false>(c, inst, a.y0, nq + i));
        const double sc = (u0*1.0 > 1.0) ? 1.0/u0 : 1.0;
        const double v = sc*ldS<false>(c, inst, a.ys, nq + i);
        uAcc = normAcc(uAcc, v, inf);
    }
    for (int b = 1 + threadIdx.x; b < c.nb; b += LP_THREADS) {
        const BodyConst& bc = c.bodies[b];
        int first = 0;
        if (bc.joint == JT_BALL || bc.joint == JT_FREE) {
            double q[4], e[4], o[4];
            for (int i = 0; i < 4; ++i) { q[i] = ldS<false>(c, inst, a.y, bc.q0 + i); e[i] = ldS<false>(c, inst, a.ys, bc.q0 + i); }
            const V3 du = quatNInvTimes(q, e);
            quatNTimes(q, du, o);
            for (int i = 0; i < 4; ++i) qAcc = normAcc(qAcc, o[i], inf);
            first = 4;
        }
        const int nqb = nqOfJoint(bc.joint);
        for (int i = first; i < nqb; ++i) qAcc = normAcc(qAcc, ldS<false>(c, inst, a.ys, bc.q0 + i), inf);
    }
    uAcc = blockReduce(uAcc, inf, red); qAcc = blockReduce(qAcc, inf, red);
    const double qNorm = inf ? qAcc : (nq ? sqrt(qAcc/nq) : 0.0), uNorm = inf ? uAcc : (nu ? sqrt(uAcc/nu) : 0.0);
    return normMax(uNorm, qNorm);
}

template <int OP>
__global__ void __launch_bounds__(LP_THREADS, 2) lpKernel(const KArgs a) {
    __shared__ double red[LP_THREADS/32];
    __shared__ Ctx sctx;
    const int inst = blockIdx.x;
    if (threadIdx.x == 0) fillCtx(sctx, a, a.tables, OP == OP_RKM);
    __syncthreads();
    const Ctx& c = sctx;
    LpLevels L; L.order = reinterpret_cast<const int*>(a.tables + a.levelOrderOff);
    L.start = reinterpret_cast<const int*>(a.tables + a.levelStartOff); L.nlevels = a.nlevels;

    if constexpr (OP == OP_KIN) {
        lpOutward(L, [&](int b) { kinDispatch(c, b, inst, c.qdot); });
    } else if constexpr (OP == OP_ABI) {
        lpInward(L, [&](int b) { inwardDispatch<IN_ABI>(c, b, inst); });
    } else if constexpr (OP == OP_EVAL) {
        lpEval(c, inst, L, c.qdot, c.udot, c.qdotdot);
    } else if constexpr (OP == OP_CALCACC) {
        lpInward(L,  [&](int b) { inwardDispatch<IN_Z | IN_BIAS>(c, b, inst); });
        lpOutward(L, [&](int b) { outwardDispatch<true>(c, b, inst, c.vecOut, nullptr); });
    } else if constexpr (OP == OP_MULM) {
        lpOutward(L, [&](int b) { idOutDispatch<false>(c, b, inst); });
        lpInward(L,  [&](int b) { idInDispatch<false>(c, b, inst); });
    } else if constexpr (OP == OP_MULMINV) {
        lpInward(L,  [&](int b) { inwardDispatch<IN_Z>(c, b, inst); });
        lpOutward(L, [&](int b) { outwardDispatch<false>(c, b, inst, c.vecOut, nullptr); });
    } else if constexpr (OP == OP_RESID) {
        lpOutward(L, [&](int b) { idOutDispatch<true>(c, b, inst); });
        lpInward(L,  [&](int b) { idInDispatch<true>(c, b, inst); });
    } else if constexpr (OP == OP_RKM) {
        const int nq = c.nq, ny = c.nq + c.nu; const long long uoff = (long long)nq*c.sStride;
        const double h = a.h; double err = 0; int nproj = 0;
        for (int s = 0; s < a.nsteps; ++s) {
            lpEval(c, inst, L, a.f0, a.f0 + uoff, nullptr);
            for (int i = threadIdx.x; i < ny; i += LP_THREADS) { const double y0 = ldS<false>(c, inst, a.y, i); stS<false>(c, inst, a.y0, i, y0); stS<false>(c, inst, a.y, i, y0 + (h/3)*ldS<false>(c, inst, a.f0, i)); }
            __syncthreads();
            lpEval(c, inst, L, a.fa, a.fa + uoff, nullptr);
            for (int i = threadIdx.x; i < ny; i += LP_THREADS) stS<false>(c, inst, a.y, i, ldS<false>(c, inst, a.y0, i) + (h/6)*(ldS<false>(c, inst, a.f0, i) + ldS<false>(c, inst, a.fa, i)));
            __syncthreads();
            lpEval(c, inst, L, a.fa, a.fa + uoff, nullptr);
            for (int i = threadIdx.x; i < ny; i += LP_THREADS) stS<false>(c, inst, a.y, i, ldS<false>(c, inst, a.y0, i) + (h/8)*(ldS<false>(c, inst, a.f0, i) + 3*ldS<false>(c, inst, a.fa, i)));
            __syncthreads();
            lpEval(c, inst, L, a.fb, a.fb + uoff, nullptr);
            for (int i = threadIdx.x; i < ny; i += LP_THREADS) {
                const double ys = ldS<false>(c, inst, a.y0, i) + (h/2)*(ldS<false>(c, inst, a.f0, i) - 3*ldS<false>(c, inst, a.fa, i) + 4*ldS<false>(c, inst, a.fb, i));
                stS<false>(c, inst, a.ys, i, ys); stS<false>(c, inst, a.y, i, ys);
            }
            __syncthreads();
            lpEval(c, inst, L, a.fa, a.fa + uoff, nullptr);
            for (int i = threadIdx.x; i < ny; i += LP_THREADS) {
                const double y1 = ldS<false>(c, inst, a.y0, i) + (h/6)*(ldS<false>(c, inst, a.f0, i) + 4*ldS<false>(c, inst, a.fb, i) + ldS<false>(c, inst, a.fa, i));
                stS<false>(c, inst, a.y, i, y1); stS<false>(c, inst, a.ys, i, 0.2*fabs(y1 - ldS<false>(c, inst, a.ys, i)));
            }
            __syncthreads();
            err = lpErrorNorm(c, inst, a, red);
            if (c.nquat > 0 && !(err > 16.0*a.accuracy)) {       // uniform across the CTA
                double acc = 0; const bool inf = a.useInfNorm != 0;
                for (int b = 1 + threadIdx.x; b < c.nb; b += LP_THREADS) {
                    const BodyConst& bc = c.bodies[b];
                    if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                    double n2 = 0; for (int i = 0; i < 4; ++i) { const double qi = ldS<false>(c, inst, a.y, bc.q0 + i); n2 += qi*qi; }
                    const double e = sqrt(n2) - 1.0;
                    acc = normAcc(acc, e, inf);
                }
                acc = blockReduce(acc, inf, red);
                const double quatNorm = inf ? acc : sqrt(acc/c.nquat);
                if (quatNorm > projectionLimit(a.consTol)) err = CUDART_INF;      // convergence failure (AbstractIntegratorRep.cpp:165-190)
                else if (quatNorm > a.consTol || a.projectEveryStep) {
                    for (int b = 1 + threadIdx.x; b < c.nb; b += LP_THREADS) {
                        const BodyConst& bc = c.bodies[b];
                        if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                        double q[4], e[4], n2 = 0;
                        for (int i = 0; i < 4; ++i) { q[i] = ldS<false>(c, inst, a.y, bc.q0 + i); e[i] = ldS<false>(c, inst, a.ys, bc.q0 + i); n2 += q[i]*q[i]; }
                        const double n = sqrt(n2); double dt = 0;
                        for (int i = 0; i < 4; ++i) { q[i] = q[i]/n; dt += e[i]*q[i]; }
                        for (int i = 0; i < 4; ++i) { stS<false>(c, inst, a.y, bc.q0 + i, q[i]); stS<false>(c, inst, a.ys, bc.q0 + i, e[i] - dt*q[i]); }
                    }
                    __syncthreads();
                    ++nproj;
                    err = lpErrorNorm(c, inst, a, red);
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            a.tcur[inst] += a.nsteps*a.h; a.errNorm[inst] = err; a.projCount[inst] += nproj;
            if (a.status && !finiteNorm(err)) a.status[inst] |= 1;
        }
    }
}
template <int OP> cudaError_t launchLpOp(const KArgs& a, cudaStream_t stream) {
    lpKernel<OP><<<a.N, LP_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

//==============================================================================================
// Grid-level-parallel integrator (plan 4): wide trees, small batches.
// A work item is (one body of a tree level, one warp of 32 instances): every warp of a persistent,
// co-resident grid takes items of the current level in a grid-stride loop, so a level of w bodies
// offers w * N/32 warps of work to the whole GPU (plan 3 gives one CTA per instance and idles most
// of its threads on narrow levels).  Lanes are instances, so the joint switch stays warp-uniform
// and the CTA-blocked records are read fully coalesced.  Levels are separated by a grid barrier
// (a monotonic counter; the grid is sized by the occupancy API so every CTA is resident).
//==============================================================================================
constexpr int GL_THREADS = 128;

__device__ __forceinline__ void gridBarrier(unsigned long long* bar, unsigned nblocks, unsigned long long& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;                   // 64-bit: 141 barriers per step never wrap
        __threadfence();
        atomicAdd(bar, 1ULL);
        unsigned long long v;
        do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory"); } while (v < target);
    }
    __syncthreads();
}
// f(body, instance) for every (body of level l, instance); lanes of a warp = 32 consecutive instances
template <class F> __device__ __forceinline__ void glLevel(const LpLevels& L, int l, int N, F f) {
    const int nwi = (N + 31) >> 5;                                  // instance warps
    const int items = (L.start[l+1] - L.start[l])*nwi;
    const int warp = (blockIdx.x*GL_THREADS + threadIdx.x) >> 5, nwarps = (gridDim.x*GL_THREADS) >> 5, lane = threadIdx.x & 31;
    for (int it = warp; it < items; it += nwarps) {
        const int body = L.order[L.start[l] + it / nwi], inst = (it % nwi)*32 + lane;
        if (inst < N) f(body, inst);
    }
}
__device__ __forceinline__ double warpReduce(double v, bool isMax) {
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = isMax ? normMax(v, t) : v + t; }
    return v;
}
// IntegratorRep::calcErrorNorm for one instance by one warp (lanes over slots / bodies), cf. lpErrorNorm
__device__ double glErrorNorm(const Ctx& c, const int inst, const KArgs& a) {
    const int nq = c.nq, nu = c.nu, lane = threadIdx.x & 31; const bool inf = a.useInfNorm != 0;
    double uAcc = 0, qAcc = 0;
    for (int i = lane; i < nu; i += 32) {
        const double u0 = fabs(ldS<false>(c, inst, a.y0, nq + i));
        const double sc = (u0*1.0 > 1.0) ? 1.0/u0 : 1.0;
        const double v = sc*ldS<false>(c, inst, a.ys, nq + i);
        uAcc = normAcc(uAcc, v, inf);
    }
    for (int b = 1 + lane; b < c.nb; b += 32) {
        const BodyConst& bc = c.bodies[b];
        int first = 0;
        if (bc.joint == JT_BALL || bc.joint == JT_FREE) {
            double q[4], e[4], o[4];
            for (int i = 0; i < 4; ++i) { q[i] = ldS<false>(c, inst, a.y, bc.q0 + i); e[i] = ldS<false>(c, inst, a.ys, bc.q0 + i); }
            const V3 du = quatNInvTimes(q, e);
            quatNTimes(q, du, o);
            for (int i = 0; i < 4; ++i) qAcc = normAcc(qAcc, o[i], inf);
            first = 4;
        }
        const int nqb = nqOfJoint(bc.joint);
        for (int i = first; i < nqb; ++i) qAcc = normAcc(qAcc, ldS<false>(c, inst, a.ys, bc.q0 + i), inf);
    }
    uAcc = warpReduce(uAcc, inf); qAcc = warpReduce(qAcc, inf);
    const double qNorm = inf ? qAcc : (nq ? sqrt(qAcc/nq) : 0.0), uNorm = inf ? uAcc : (nu ? sqrt(uAcc/nu) : 0.0);
    return normMax(uNorm, qNorm);
}

#ifndef SBK_GL_MINB
#define SBK_GL_MINB 2
#endif
__global__ void __launch_bounds__(GL_THREADS, SBK_GL_MINB) glRkmKernel(const KArgs a) {
    __shared__ Ctx sctx;
    if (threadIdx.x == 0) fillCtx(sctx, a, a.tables, true);
    __syncthreads();
    const Ctx& c = sctx;
    LpLevels L; L.order = reinterpret_cast<const int*>(a.tables + a.levelOrderOff);
    L.start = reinterpret_cast<const int*>(a.tables + a.levelStartOff); L.nlevels = a.nlevels;
    unsigned long long target = 0; unsigned long long* bar = reinterpret_cast<unsigned long long*>(a.taskCounter);
    const int N = a.N, nq = c.nq, ny = c.nq + c.nu; const long long uoff = (long long)nq*N;
    const long long tid = (long long)blockIdx.x*GL_THREADS + threadIdx.x, nth = (long long)gridDim.x*GL_THREADS, nel = (long long)ny*N;
    const double h = a.h;
    for (int s = 0; s < a.nsteps; ++s) {
#pragma unroll 1
        for (int stage = 0; stage < 5; ++stage) {
            double* fdst = stage == 0 ? a.f0 : (stage == 3 ? a.fb : a.fa);
            double* qd = fdst; double* ud = fdst + uoff;
#pragma unroll 1
            for (int l = 1; l < L.nlevels; ++l) { glLevel(L, l, N, [&](int b, int i) { kinDispatch<true>(c, b, i, qd); }); gridBarrier(bar, gridDim.x, target); }
            if (c.ntp) { for (long long i = tid; i < N; i += nth) twoPointPass<true>(c, (int)i); gridBarrier(bar, gridDim.x, target); }
#pragma unroll 1
            for (int l = L.nlevels - 1; l >= 1; --l) { glLevel(L, l, N, [&](int b, int i) { inwardDispatch<IN_ABI | IN_Z | IN_BIAS | IN_FORCES, true>(c, b, i); }); gridBarrier(bar, gridDim.x, target); }
#pragma unroll 1
            for (int l = 1; l < L.nlevels; ++l) { glLevel(L, l, N, [&](int b, int i) { outwardDispatch<true, true>(c, b, i, ud, nullptr); }); gridBarrier(bar, gridDim.x, target); }
            // stage combination over the flat [slot][N] vectors (element e = slot*N + instance)
            for (long long e = tid; e < nel; e += nth) {
                const double y0 = stage == 0 ? a.y[e] : a.y0[e], f0 = a.f0[e];
                if (stage == 0)      { a.y0[e] = y0; a.y[e] = y0 + (h/3)*f0; }
                else if (stage == 1) { a.y[e] = y0 + (h/6)*(f0 + a.fa[e]); }
                else if (stage == 2) { a.y[e] = y0 + (h/8)*(f0 + 3*a.fa[e]); }
                else if (stage == 3) { const double ys = y0 + (h/2)*(f0 - 3*a.fa[e] + 4*a.fb[e]); a.ys[e] = ys; a.y[e] = ys; }
                else                 { const double y1 = y0 + (h/6)*(f0 + 4*a.fb[e] + a.fa[e]); a.y[e] = y1; a.ys[e] = 0.2*fabs(y1 - a.ys[e]); }
            }
            gridBarrier(bar, gridDim.x, target);
        }
        // error norm and quaternion projection: one warp per instance (AbstractIntegratorRep.cpp:137-208)
        {
            const int warp = (int)(tid >> 5), nwarps = (int)(nth >> 5), lane = threadIdx.x & 31;
            for (int inst = warp; inst < N; inst += nwarps) {
                double err = glErrorNorm(c, inst, a); int proj = 0;
                if (c.nquat > 0 && !(err > 16.0*a.accuracy)) {
                    double acc = 0; const bool inf = a.useInfNorm != 0;
                    for (int b = 1 + lane; b < c.nb; b += 32) {
                        const BodyConst& bc = c.bodies[b];
                        if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                        double n2 = 0; for (int i = 0; i < 4; ++i) { const double qi = ldS<false>(c, inst, a.y, bc.q0 + i); n2 += qi*qi; }
                        const double e = sqrt(n2) - 1.0;
                        acc = normAcc(acc, e, inf);
                    }
                    acc = warpReduce(acc, inf);
                    const double quatNorm = inf ? acc : sqrt(acc/c.nquat);
                    if (quatNorm > projectionLimit(a.consTol)) err = CUDART_INF;
                    else if (quatNorm > a.consTol || a.projectEveryStep) {
                        for (int b = 1 + lane; b < c.nb; b += 32) {
                            const BodyConst& bc = c.bodies[b];
                            if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                            double q[4], e[4], n2 = 0;
                            for (int i = 0; i < 4; ++i) { q[i] = ldS<false>(c, inst, a.y, bc.q0 + i); e[i] = ldS<false>(c, inst, a.ys, bc.q0 + i); n2 += q[i]*q[i]; }
                            const double n = sqrt(n2); double dt = 0;
                            for (int i = 0; i < 4; ++i) { q[i] = q[i]/n; dt += e[i]*q[i]; }
                            for (int i = 0; i < 4; ++i) { stS<false>(c, inst, a.y, bc.q0 + i, q[i]); stS<false>(c, inst, a.ys, bc.q0 + i, e[i] - dt*q[i]); }
                        }
                        __syncwarp();
                        proj = 1;
                        err = glErrorNorm(c, inst, a);
                    }
                }
                if (lane == 0) {
                    a.tcur[inst] += a.h; a.errNorm[inst] = err; a.projCount[inst] += proj;
                    if (a.status && !finiteNorm(err)) a.status[inst] |= 1;
                }
            }
        }
        gridBarrier(bar, gridDim.x, target);
    }
}
// API operations of plan 4 with the same mapping (one launch = the sweeps of one operation).
template <int OP>
__global__ void __launch_bounds__(GL_THREADS, SBK_GL_MINB) glOpKernel(const KArgs a) {
    __shared__ Ctx sctx;
    if (threadIdx.x == 0) fillCtx(sctx, a, a.tables, false);
    __syncthreads();
    const Ctx& c = sctx;
    LpLevels L; L.order = reinterpret_cast<const int*>(a.tables + a.levelOrderOff);
    L.start = reinterpret_cast<const int*>(a.tables + a.levelStartOff); L.nlevels = a.nlevels;
    unsigned long long target = 0; unsigned long long* bar = reinterpret_cast<unsigned long long*>(a.taskCounter);
    const int N = a.N;
    auto outward = [&](auto f) { for (int l = 1; l < L.nlevels; ++l) { glLevel(L, l, N, f); gridBarrier(bar, gridDim.x, target); } };
    auto inward  = [&](auto f) { for (int l = L.nlevels - 1; l >= 1; --l) { glLevel(L, l, N, f); gridBarrier(bar, gridDim.x, target); } };
    if constexpr (OP == OP_KIN) {
        outward([&](int b, int i) { kinDispatch<true>(c, b, i, c.qdot); });
    } else if constexpr (OP == OP_ABI) {
        inward([&](int b, int i) { inwardDispatch<IN_ABI, true>(c, b, i); });
    } else if constexpr (OP == OP_EVAL) {
        outward([&](int b, int i) { kinDispatch<true>(c, b, i, c.qdot); });
        if (c.ntp) {
            const long long tid = (long long)blockIdx.x*GL_THREADS + threadIdx.x, nth = (long long)gridDim.x*GL_THREADS;
            for (long long i = tid; i < N; i += nth) twoPointPass<true>(c, (int)i);
            gridBarrier(bar, gridDim.x, target);
        }
        inward([&](int b, int i) { inwardDispatch<IN_ABI | IN_Z | IN_BIAS | IN_FORCES, true>(c, b, i); });
        outward([&](int b, int i) { outwardDispatch<true, true>(c, b, i, c.udot, c.qdotdot); });
    } else if constexpr (OP == OP_CALCACC) {
        inward([&](int b, int i) { inwardDispatch<IN_Z | IN_BIAS, true>(c, b, i); });
        outward([&](int b, int i) { outwardDispatch<true, true>(c, b, i, c.vecOut, nullptr); });
    } else if constexpr (OP == OP_MULM) {
        outward([&](int b, int i) { idOutDispatch<false, true>(c, b, i); });
        inward([&](int b, int i) { idInDispatch<false, true>(c, b, i); });
    } else if constexpr (OP == OP_MULMINV) {
        inward([&](int b, int i) { inwardDispatch<IN_Z, true>(c, b, i); });
        outward([&](int b, int i) { outwardDispatch<false, true>(c, b, i, c.vecOut, nullptr); });
    } else if constexpr (OP == OP_RESID) {
        outward([&](int b, int i) { idOutDispatch<true, true>(c, b, i); });
        inward([&](int b, int i) { idInDispatch<true, true>(c, b, i); });
    }
}
template <class K> cudaError_t launchGlCoop(K kernel, const KArgs& a, cudaStream_t stream) {
    int dev = 0, sms = 0, perSm = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, GL_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (perSm < 1) return cudaErrorLaunchOutOfResources;
    e = cudaMemsetAsync(a.taskCounter, 0, 2*sizeof(int), stream);         // the 64-bit barrier counter
    if (e != cudaSuccess) return e;
    void* args[] = { const_cast<KArgs*>(&a) };
    // cooperative launch: the driver guarantees (or refuses) co-residency of the whole grid
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3(sms*perSm), dim3(GL_THREADS), args, 0, stream);
}
cudaError_t launchGlOpImpl(KernelOp op, const KArgs& a, cudaStream_t stream) {
    switch (op) {
        case OP_KIN:     return launchGlCoop(glOpKernel<OP_KIN>, a, stream);
        case OP_ABI:     return launchGlCoop(glOpKernel<OP_ABI>, a, stream);
        case OP_EVAL:    return launchGlCoop(glOpKernel<OP_EVAL>, a, stream);
        case OP_CALCACC: return launchGlCoop(glOpKernel<OP_CALCACC>, a, stream);
        case OP_MULM:    return launchGlCoop(glOpKernel<OP_MULM>, a, stream);
        case OP_MULMINV: return launchGlCoop(glOpKernel<OP_MULMINV>, a, stream);
        case OP_RESID:   return launchGlCoop(glOpKernel<OP_RESID>, a, stream);
        default: break;
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchGlRkmImpl(const KArgs& a, cudaStream_t stream) { return launchGlCoop(glRkmKernel, a, stream); }

__device__ __forceinline__ long long instOffsetK(const KArgs& a, int k) {
    return (long long)(k >> a.cShift)*a.cSpan + (long long)(k & a.cMask)*a.cInstStride;
}
__global__ void initGroundKernel(const KArgs a) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    double* rec = a.cache + instOffsetK(a, k);     // Ground is record base 0 in every plan
    for (int f = 0; f < F_H; ++f) rec[(long long)f*a.cStride] = (f == F_XGB || f == F_XGB+4 || f == F_XGB+8) ? 1.0 : 0.0;
}

// Integrator::initialize (Integrator.cpp:367-377, ProjectOptions::ForceProjection): SimbodyMatterSubsystemRep::normalizeQuaternions
// (:4448-4476) -> RBNodeBall/Free::enforceQuaternionConstraints (RigidBodyNodeSpec_Ball.h:417-434): quat /= |quat|, one projection counted.
__global__ void initProjectKernel(const KArgs a) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    for (int b = 1; b < a.nb; ++b) {
        if (bodies[b].joint != JT_BALL && bodies[b].joint != JT_FREE) continue;
        double* q = a.y + (long long)bodies[b].q0*a.N + k; const long long N = a.N;
        const double n = sqrt(q[0]*q[0] + q[N]*q[N] + q[2*N]*q[2*N] + q[3*N]*q[3*N]);
        q[0] = q[0]/n; q[N] = q[N]/n; q[2*N] = q[2*N]/n; q[3*N] = q[3*N]/n;
    }
    a.projCount[k] += 1;
}

// 32x32 tiled transpose: src is [rows][cols] row-major -> dst [cols][rows]
__global__ void transposeKernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int cols) {
    __shared__ double tile[32][33];
    int c = blockIdx.x*32 + threadIdx.x, r0 = blockIdx.y*32;
    for (int j = threadIdx.y; j < 32; j += 8) { const int r = r0 + j; if (r < rows && c < cols) tile[j][threadIdx.x] = src[(long long)r*cols + c]; }
    __syncthreads();
    const int r = r0 + threadIdx.x; int c0 = blockIdx.x*32;
    for (int j = threadIdx.y; j < 32; j += 8) { c = c0 + j; if (r < rows && c < cols) dst[(long long)c*rows + r] = tile[threadIdx.x][j]; }
}

__global__ void gatherBodyFieldKernel(const KArgs a, int fieldOffset, int width, double* out) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    for (int b = 0; b < a.nb; ++b) {
        const double* rec = a.cache + bodies[b].cacheBase + instOffsetK(a, k);
        for (int i = 0; i < width; ++i) out[((long long)b*width + i)*a.N + k] = rec[(long long)(fieldOffset + i)*a.cStride];
    }
}

// Kinetic and potential energy of every instance from the realized position/velocity records:
//   KE = sum_b 1/2 V_GB . (Mk_G V_GB) in level order   (SimbodyMatterSubsystemRep.cpp:5234-5246, RigidBodyNode.cpp:182-188)
//   PE = gravity: -m (g . p_G_CB) per body (Force_Gravity.cpp:555) + springs: k (q-q0)^2 / 2 (Force.cpp:354-361)
__global__ void energyKernel(const KArgs a, double* ke, double* pe) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    const ForceConst* forces = reinterpret_cast<const ForceConst*>(a.tables + a.forcesOff);
    const int* order = reinterpret_cast<const int*>(a.tables + a.levelOrderOff);
    double kin = 0, pot = 0;
    for (int i = 1; i < a.nb; ++i) {                       // level order (order[0] is Ground)
        const BodyConst& bc = bodies[order[i]];
        CacheRef me; me.p = a.cache + bc.cacheBase + instOffsetK(a, k); me.stride = a.cStride;
        const SV V = me.ldSV(F_VGB);
        kin += dot(V, mulSpatialInertia(bc.mass, me.ld3(F_MK), me.ldS3(F_MK + 3), V))/2;
    }
    for (int b = 1; b < a.nb; ++b) {                       // MobilizedBodyIndex order
        const BodyConst& bc = bodies[b];
        CacheRef me; me.p = a.cache + bc.cacheBase + instOffsetK(a, k); me.stride = a.cStride;
        const V3 pc = me.ld3(F_XGB + 9) + me.ld3(F_MK);    // p_G_CB = p_GB + R_GB*com_B
        pot -= bc.mass*(a.gx*pc.x + a.gy*pc.y + a.gz*pc.z + 0.0);
    }
    for (int b = 1; b < a.nb; ++b) {
        const BodyConst& bc = bodies[b];
        for (int f = 0; f < bc.nforce; ++f) {
            const ForceConst fc = forces[bc.forceStart + f];
            if (fc.kind != FK_SPRING) continue;
            const double dq = a.y[(long long)(bc.q0 + fc.coord)*a.N + k] - fc.b;
            pot += fc.a*(dq*dq)/2;
        }
    }
    const TwoPointConst* tp = reinterpret_cast<const TwoPointConst*>(a.tables + a.tpOff);
    for (int e = 0; e < a.ntp; ++e) {                      // TwoPointLinearSpringImpl::calcPotentialEnergy (Force.cpp:122-137)
        if (tp[e].kind != TP_SPRING) continue;
        CacheRef r1, r2; r1.p = a.cache + tp[e].cacheBase1 + instOffsetK(a, k); r1.stride = a.cStride; r2.p = a.cache + tp[e].cacheBase2 + instOffsetK(a, k); r2.stride = a.cStride;
        const V3 p1 = r1.ld3(F_XGB + 9) + mul(r1.ldM3(F_XGB), mk(tp[e].s1[0], tp[e].s1[1], tp[e].s1[2]));
        const V3 p2 = r2.ld3(F_XGB + 9) + mul(r2.ldM3(F_XGB), mk(tp[e].s2[0], tp[e].s2[1], tp[e].s2[2]));
        const V3 r = p2 - p1; const double stretch = sqrt(dot(r, r)) - tp[e].b;
        pot += tp[e].a*stretch*stretch/2;
    }
    if (ke) ke[k] = kin;
    if (pe) pe[k] = pot;
}

// calcMobilizerReactionForces (SimbodyMatterSubsystemRep.cpp:5788-5832) from the records of a realized
// acceleration stage: FB = zPlus + PPlus*(~Phi A_GP) at the body origin, reported at the origin of the
// outboard frame M (FM = (m - (R_GB p_BM) x f, f)), expressed in Ground.  Ground's entry collects the
// base bodies (RigidBodyNode_Weld.cpp:197-222; the lowered force elements apply nothing to Ground).
__global__ void reactionKernel(const KArgs a, double* out) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    auto rec = [&](int b) { CacheRef r; r.p = a.cache + bodies[b].cacheBase + instOffsetK(a, k); r.stride = a.cStride; return r; };
    auto put = [&](int b, SV F) {
        double* o = out + (long long)b*6*a.N + k;
        o[0] = F.w.x; o[(long long)a.N] = F.w.y; o[2LL*a.N] = F.w.z; o[3LL*a.N] = F.v.x; o[4LL*a.N] = F.v.y; o[5LL*a.N] = F.v.z;
    };
    SV z0 = zeroSV();
    for (int b = 1; b < a.nb; ++b) {
        const BodyConst& bc = bodies[b];
        const CacheRef me = rec(b), pa = rec(bc.parent);
        const V3 pPB = me.ld3(F_XGB + 9) - pa.ld3(F_XGB + 9);
        const SV APlus = phiT(pPB, pa.ldSV(F_AGB));
        const SV zP = me.ldSV(F_ZPLUS);
        const SV FB = zP + mul(me.ldABI(F_PPLUS), APlus);
        const V3 pBM = mul(me.ldM3(F_XGB), mk(bc.p_BM[0], bc.p_BM[1], bc.p_BM[2]));
        SV FM; FM.w = FB.w - cross(pBM, FB.v); FM.v = FB.v;
        put(b, FM);
        if (bc.parent == 0) z0 = z0 + phi(me.ld3(F_L), zP);
    }
    if (a.ntp) {        // a two-point element may act on Ground: z[0] = -F[0] + sum Phi zPlus (RigidBodyNode_Weld.cpp:213-222)
        z0.w = z0.w - mk(a.f2[k], a.f2[(long long)a.N + k], a.f2[2LL*a.N + k]); z0.v = z0.v - mk(a.f2[3LL*a.N + k], a.f2[4LL*a.N + k], a.f2[5LL*a.N + k]);
    }
    put(0, z0);
}

// multiplyBySystemJacobian (RigidBodyNodeSpec.cpp:760-780): Jv_b = ~Phi_b Jv_parent + H_b v_b, base to tip.
__global__ void jacobianKernel(const KArgs a, const double* v, double* out) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    const long long N = a.N;
    for (int i = 0; i < 6; ++i) out[i*N + k] = 0.0;
    for (int b = 1; b < a.nb; ++b) {
        const BodyConst& bc = bodies[b];
        CacheRef me; me.p = a.cache + bc.cacheBase + instOffsetK(a, k); me.stride = a.cStride;
        const double* po = out + (long long)bc.parent*6*N + k;
        SV JP; JP.w = mk(po[0], po[N], po[2*N]); JP.v = mk(po[3*N], po[4*N], po[5*N]);
        const SV sh = phiT(me.ld3(F_L), JP);
        SV Hv = zeroSV();
        const int d = dofOfJoint(bc.joint);
        for (int j = 0; j < d; ++j) Hv = Hv + v[(long long)(bc.u0 + j)*N + k]*me.ldSV(F_H + 6*j);
        const SV J = sh + Hv;
        double* o = out + (long long)b*6*N + k;
        o[0] = J.w.x; o[N] = J.w.y; o[2*N] = J.w.z; o[3*N] = J.v.x; o[4*N] = J.v.y; o[5*N] = J.v.z;
    }
}

// multiplyBySystemJacobianTranspose (RigidBodyNodeSpec.cpp:790-815): z_b = F_b + sum Phi_c z_c, out_b = ~H_b z_b,
// tip to base.  z [nb*6][N] is scratch.
__global__ void jacobianTransposeKernel(const KArgs a, const double* F, double* z, double* out) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    const int* children = reinterpret_cast<const int*>(a.tables + a.childrenOff);
    const long long N = a.N;
    auto ld6 = [&](const double* p, int b) { const double* q = p + (long long)b*6*N + k; SV r; r.w = mk(q[0], q[N], q[2*N]); r.v = mk(q[3*N], q[4*N], q[5*N]); return r; };
    for (int b = a.nb - 1; b >= 1; --b) {
        const BodyConst& bc = bodies[b];
        CacheRef me; me.p = a.cache + bc.cacheBase + instOffsetK(a, k); me.stride = a.cStride;
        SV zb = ld6(F, b);
        for (int c = 0; c < bc.nchild; ++c) {
            const int cb = children[bc.childStart + c];
            CacheRef ch; ch.p = a.cache + bodies[cb].cacheBase + instOffsetK(a, k); ch.stride = a.cStride;
            zb = zb + phi(ch.ld3(F_L), ld6(z, cb));
        }
        double* o = z + (long long)b*6*N + k;
        o[0] = zb.w.x; o[N] = zb.w.y; o[2*N] = zb.w.z; o[3*N] = zb.v.x; o[4*N] = zb.v.y; o[5*N] = zb.v.z;
        const int d = dofOfJoint(bc.joint);
        for (int j = 0; j < d; ++j) out[(long long)(bc.u0 + j)*N + k] = dot(me.ldSV(F_H + 6*j), zb);
    }
}

// calcCompositeBodyInertias (SimbodyMatterSubsystemRep.cpp:5196-5205, RigidBodyNode.cpp:231-243): the spatial inertia
// of the rigid body made by locking every joint outboard of a body, about the body origin, in Ground:
//   R_b = Mk_b + sum_children R_c.shift(-l_c)     (SpatialInertia shift / +=, MassProperties.h:1037-1045,1110-1116)
// out [nb*10][N]: mass, com (3), unit inertia xx yy zz xy xz yz.  Ground: infinite mass (RigidBodyNode_Weld.cpp:168-171).
__global__ void cbiKernel(const KArgs a, double* out) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const BodyConst* bodies = reinterpret_cast<const BodyConst*>(a.tables);
    const int* children = reinterpret_cast<const int*>(a.tables + a.childrenOff);
    const long long N = a.N;
    auto row = [&](int b, int i) -> double& { return out[((long long)b*10 + i)*N + k]; };
    row(0, 0) = CUDART_INF; row(0, 1) = 0; row(0, 2) = 0; row(0, 3) = 0; row(0, 4) = 1; row(0, 5) = 1; row(0, 6) = 1; row(0, 7) = 0; row(0, 8) = 0; row(0, 9) = 0;
    for (int b = a.nb - 1; b >= 1; --b) {
        const BodyConst& bc = bodies[b];
        CacheRef me; me.p = a.cache + bc.cacheBase + instOffsetK(a, k); me.stride = a.cStride;
        double m = bc.mass; V3 p = me.ld3(F_MK); S3 G = me.ldS3(F_MK + 3);
        for (int c = 0; c < bc.nchild; ++c) {
            const int cb = children[bc.childStart + c];
            CacheRef ch; ch.p = a.cache + bodies[cb].cacheBase + instOffsetK(a, k); ch.stride = a.cStride;
            const double mc = row(cb, 0); const V3 pc = mk(row(cb, 1), row(cb, 2), row(cb, 3));
            S3 Gc; Gc.xx = row(cb, 4); Gc.yy = row(cb, 5); Gc.zz = row(cb, 6); Gc.xy = row(cb, 7); Gc.xz = row(cb, 8); Gc.yz = row(cb, 9);
            // shift(-l): to the centroid, then out to the parent origin (the com is at pc + l from there)
            Gc.xx -= pc.y*pc.y + pc.z*pc.z; Gc.yy -= pc.x*pc.x + pc.z*pc.z; Gc.zz -= pc.x*pc.x + pc.y*pc.y;
            Gc.xy -= -pc.x*pc.y; Gc.xz -= -pc.x*pc.z; Gc.yz -= -pc.y*pc.z;
            const V3 pn = pc + ch.ld3(F_L);
            Gc.xx += pn.y*pn.y + pn.z*pn.z; Gc.yy += pn.x*pn.x + pn.z*pn.z; Gc.zz += pn.x*pn.x + pn.y*pn.y;
            Gc.xy += -pn.x*pn.y; Gc.xz += -pn.x*pn.z; Gc.yz += -pn.y*pn.z;
            const double mt = m + mc, oo = 1.0/mt;
            p = oo*(m*p + mc*pn);
            G.xx = oo*(m*G.xx + mc*Gc.xx); G.yy = oo*(m*G.yy + mc*Gc.yy); G.zz = oo*(m*G.zz + mc*Gc.zz);
            G.xy = oo*(m*G.xy + mc*Gc.xy); G.xz = oo*(m*G.xz + mc*Gc.xz); G.yz = oo*(m*G.yz + mc*Gc.yz);
            m = mt;
        }
        row(b, 0) = m; row(b, 1) = p.x; row(b, 2) = p.y; row(b, 3) = p.z;
        row(b, 4) = G.xx; row(b, 5) = G.yy; row(b, 6) = G.zz; row(b, 7) = G.xy; row(b, 8) = G.xz; row(b, 9) = G.yz;
    }
}

// Memory-pattern probe (bench/diagnostics only): one thread per instance walks `nb` records of
// `rowsIn` + `rowsOut` rows in the [row][N] layout of the thread-per-instance plan, loading rowsIn
// doubles and storing rowsOut doubles per record with no arithmetic to speak of.  It measures what
// the memory system delivers for the plan's access pattern at a given occupancy.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) memPatternKernel(double* buf, int N, int nb, int rowsIn, int rowsOut, int sweeps) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= N) return;
    const long long rec = (long long)(rowsIn + rowsOut)*N;
    double acc = 0;
    for (int s = 0; s < sweeps; ++s)
        for (int b = 0; b < nb; ++b) {
            double* p = buf + (long long)b*rec + k;
            double v[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) v[i] = (i < rowsIn) ? __ldcg(p + (long long)i*N) : 0.0;
#pragma unroll
            for (int i = 0; i < 48; ++i) acc += v[i];
            for (int i = 0; i < rowsOut; ++i) __stcg(p + (long long)(rowsIn + i)*N, acc + i);
        }
    if (acc == 12345.678) buf[k] = acc;
}

__global__ void dfmaProbeKernel(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x*1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3,
           a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999, b = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace

cudaError_t launchTpi(KernelOp op, const KArgs& a, cudaStream_t stream) {
    switch (op) {
        case OP_KIN:     return launchOp<OP_KIN, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_ABI:     return launchOp<OP_ABI, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_EVAL:    return launchOp<OP_EVAL, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_CALCACC: return launchOp<OP_CALCACC, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_MULM:    return launchOp<OP_MULM, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_MULMINV: return launchOp<OP_MULMINV, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_RESID:   return launchOp<OP_RESID, SBK_HEAVY_MINB, JM_ALL>(a, stream);
        case OP_RKM: case OP_RKM_ADAPT: {
            // two-point force elements need every body's ground-frame transform between the sweeps: the FULL-record grid-level
            // integrator (same record layout) steps such models; error-controlled stepping is not available for them
            if (a.ntp > 0) return op == OP_RKM ? launchGlRkmImpl(a, stream) : cudaErrorNotSupported;
            // integrator kernels: instantiated per set of mobilizer kinds present in the model (one translation unit each)
            const int m = a.jointMask;
            if (a.ltables) {            // body-frame sweeps: Pin-only models get the kernel without the other mobilizers' code
                if ((m & ~JM_PIN) == 0) {
                    // fixed steps: three 128-instance work groups per SM over one staged copy of the tables (12 warps per SM at
                    // 168 registers instead of 8 at 255: +10% on the 50-link chain) when that fits the SM's shared memory
                    const bool fits = a.lstageInSmem && a.ltableBytes + launchTpiRkmLocalPin_m3_workBytes() + 2048 <= (size_t)227*1024;
                    return (op == OP_RKM && fits) ? launchTpiRkmLocalPin_m3(op, a, stream) : launchTpiRkmLocalPin_m2(op, a, stream);
                }
                return launchTpiRkmLocal_m2(op, a, stream);
            }
            // every other model (Weld, Translation, Cylinder, Planar, Gimbal, Euler-angle mode, or SBK_NOLOCAL=1): round 1's
            // ground-frame integrator with reversible kinematics, one build for all mobilizer kinds
            return launchTpiRkmAll(op, a, stream);
        }
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchLp(KernelOp op, const KArgs& a, cudaStream_t stream) {
    switch (op) {
        case OP_KIN:     return launchLpOp<OP_KIN>(a, stream);
        case OP_ABI:     return launchLpOp<OP_ABI>(a, stream);
        case OP_EVAL:    return launchLpOp<OP_EVAL>(a, stream);
        case OP_CALCACC: return launchLpOp<OP_CALCACC>(a, stream);
        case OP_MULM:    return launchLpOp<OP_MULM>(a, stream);
        case OP_MULMINV: return launchLpOp<OP_MULMINV>(a, stream);
        case OP_RESID:   return launchLpOp<OP_RESID>(a, stream);
        case OP_RKM:     return launchLpOp<OP_RKM>(a, stream);
        default: break;   // OP_RKM_ADAPT: not available in the level-parallel plan
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchGlRkm(const KArgs& a, cudaStream_t stream) { return launchGlRkmImpl(a, stream); }
cudaError_t launchGl(KernelOp op, const KArgs& a, cudaStream_t stream) { return op == OP_RKM ? launchGlRkmImpl(a, stream) : launchGlOpImpl(op, a, stream); }
bool fusedPlanSupports(int nb, const int* joints) {
    auto simple = [](int j) { return j == JT_PIN || j == JT_SLIDER; };
    if (nb == 2) return simple(joints[1]) || joints[1] == JT_UNIVERSAL;
    if (nb == 3) return simple(joints[1]) && simple(joints[2]);
    return false;
}
cudaError_t launchFusedRkm(const KArgs& a, const int* joints, bool adaptive, cudaStream_t stream) {
    if (a.nb == 2) {
        switch (joints[1]) {
            case JT_PIN:       return launchFusedT<Chain1<JT_PIN>>(a, adaptive, stream);
            case JT_SLIDER:    return launchFusedT<Chain1<JT_SLIDER>>(a, adaptive, stream);
            case JT_UNIVERSAL: return launchFusedT<Chain1<JT_UNIVERSAL>>(a, adaptive, stream);
        }
    } else if (a.nb == 3) {
        const int k = joints[1]*10 + joints[2];
        switch (k) {
            case JT_PIN*10 + JT_PIN:       return launchFusedT<Chain2<JT_PIN, JT_PIN>>(a, adaptive, stream);
            case JT_PIN*10 + JT_SLIDER:    return launchFusedT<Chain2<JT_PIN, JT_SLIDER>>(a, adaptive, stream);
            case JT_SLIDER*10 + JT_PIN:    return launchFusedT<Chain2<JT_SLIDER, JT_PIN>>(a, adaptive, stream);
            case JT_SLIDER*10 + JT_SLIDER: return launchFusedT<Chain2<JT_SLIDER, JT_SLIDER>>(a, adaptive, stream);
        }
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchInitProject(const KArgs& a, cudaStream_t stream) {
    initProjectKernel<<<(a.N + 255)/256, 256, 0, stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launchInitGround(const KArgs& a, cudaStream_t stream) {
    initGroundKernel<<<(a.N + 255)/256, 256, 0, stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launchTranspose(const double* src, double* dst, int rows, int cols, cudaStream_t stream) {
    dim3 grid((cols + 31)/32, (rows + 31)/32), block(32, 8);
    transposeKernel<<<grid, block, 0, stream>>>(src, dst, rows, cols);
    return cudaGetLastError();
}
cudaError_t launchGatherBodyField(const KArgs& a, int fieldOffset, int width, double* out, cudaStream_t stream) {
    gatherBodyFieldKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, fieldOffset, width, out);
    return cudaGetLastError();
}
cudaError_t launchEnergy(const KArgs& a, double* ke, double* pe, cudaStream_t stream) {
    energyKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, ke, pe);
    return cudaGetLastError();
}
cudaError_t launchReaction(const KArgs& a, double* out, cudaStream_t stream) {
    reactionKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, out);
    return cudaGetLastError();
}
cudaError_t launchJacobian(const KArgs& a, const double* v, double* out, cudaStream_t stream) {
    jacobianKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, v, out);
    return cudaGetLastError();
}
cudaError_t launchJacobianTranspose(const KArgs& a, const double* F, double* z, double* out, cudaStream_t stream) {
    jacobianTransposeKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, F, z, out);
    return cudaGetLastError();
}
cudaError_t launchCompositeBodyInertias(const KArgs& a, double* out, cudaStream_t stream) {
    cbiKernel<<<(a.N + 127)/128, 128, 0, stream>>>(a, out);
    return cudaGetLastError();
}
cudaError_t launchMemPattern(double* buf, int N, int nb, int rowsIn, int rowsOut, int sweeps, int minBlocks, cudaStream_t stream) {
    const int grid = (N + 127)/128;
    if (minBlocks >= 4) memPatternKernel<4><<<grid, 128, 0, stream>>>(buf, N, nb, rowsIn, rowsOut, sweeps);
    else                memPatternKernel<2><<<grid, 128, 0, stream>>>(buf, N, nb, rowsIn, rowsOut, sweeps);
    return cudaGetLastError();
}
cudaError_t launchDfmaProbe(double* out, int iters, int blocks, int threads, cudaStream_t stream) {
    dfmaProbeKernel<<<blocks, threads, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

} // namespace sbkd
