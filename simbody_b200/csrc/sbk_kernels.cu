// sbk_kernels.cu -- sm_100a kernels of the thread-per-instance plan.
//
// Mapping: one thread = one instance; a warp = 32 instances walking the SAME body at the same
// time, so the joint-type switch is warp-uniform and every cache/state access
// cache[(record+field)*N + instance] is a fully coalesced 256-byte warp transaction.
// Batch-shared body constants are staged into shared memory once per CTA by a TMA bulk copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) and then read as warp
// broadcasts.  FP64 throughout; no tensor cores (6x6 spatial operators are not a dense
// contraction).
#include "sbk_kernels.cuh"
#include "sbk_fused.cuh"

namespace sbkd {

namespace {

#ifndef SBK_TPI_THREADS
#define SBK_TPI_THREADS 128
#endif
#ifndef SBK_TPI_MINBLOCKS
#define SBK_TPI_MINBLOCKS 4
#endif
constexpr int TPI_THREADS = SBK_TPI_THREADS;

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

__device__ __forceinline__ Ctx makeCtx(const KArgs& a, const unsigned char* tables, int inst) {
    Ctx c;
    c.bodies   = reinterpret_cast<const BodyConst*>(tables);
    c.children = reinterpret_cast<const int*>(tables + a.childrenOff);
    c.forces   = reinterpret_cast<const ForceConst*>(tables + a.forcesOff);
    c.nb = a.nb; c.nq = a.nq; c.nu = a.nu; c.nquat = a.nquat;
    c.gx = a.gx; c.gy = a.gy; c.gz = a.gz;
    c.cache = a.cache; c.cStride = a.N; c.cOff = inst;
    c.sStride = a.N; c.sOff = inst;
    c.q = a.y; c.u = a.y + (long long)a.nq*a.N;
    c.qdot = a.ydot; c.udot = a.ydot ? a.ydot + (long long)a.nq*a.N : nullptr;
    c.qdotdot = a.qdotdot; c.qerr = a.qerr;
    c.fmobIn = a.fmobIn; c.FbodyIn = a.FbodyIn; c.fmobOut = a.fmobOut; c.FbodyOut = a.FbodyOut;
    c.vecIn = a.vecIn; c.vecOut = a.vecOut;
    c.status = a.status ? a.status + inst : nullptr;
    return c;
}

template <int OP, bool STAGE>
__global__ void __launch_bounds__(TPI_THREADS, SBK_TPI_MINBLOCKS) tpiKernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    const unsigned char* tables = a.tables;
    if constexpr (STAGE) { tmaStage(smem, a.tables, a.tableBytes, &mbar); tables = smem; }
    const int inst = blockIdx.x*blockDim.x + threadIdx.x;
    if (inst >= a.N) return;
    Ctx c = makeCtx(a, tables, inst);
    Carry cy; resetCarry(cy);

    if constexpr (OP == OP_KIN) {
        tpiKinematics<false>(c, cy);
    } else if constexpr (OP == OP_ABI) {
        tpiInward<IN_ABI, false>(c, cy);
    } else if constexpr (OP == OP_EVAL) {
        tpiEvalDerivatives<false>(c, cy);
    } else if constexpr (OP == OP_CALCACC) {
        tpiInward<IN_Z | IN_BIAS, false>(c, cy);
        tpiOutward<true, false>(c, cy, c.vecOut, nullptr);
    } else if constexpr (OP == OP_MULM) {
        for (int b = 1; b < c.nb; ++b) idOutDispatch<false>(c, b);
        for (int b = c.nb - 1; b >= 1; --b) idInDispatch<false>(c, b);
    } else if constexpr (OP == OP_MULMINV) {
        c.fmobIn = a.vecIn; c.FbodyIn = nullptr;
        tpiInward<IN_Z, false>(c, cy);
        tpiOutward<false, false>(c, cy, c.vecOut, nullptr);
    } else if constexpr (OP == OP_RESID) {
        for (int b = 1; b < c.nb; ++b) idOutDispatch<true>(c, b);
        for (int b = c.nb - 1; b >= 1; --b) idInDispatch<true>(c, b);
    } else if constexpr (OP == OP_RKM) {
        RkmWork w;
        w.y = a.y; w.y0 = a.y0; w.f0 = a.f0; w.fa = a.fa; w.fb = a.fb; w.ys = a.ys;
        w.accuracy = a.accuracy; w.consTol = a.consTol; w.useInfNorm = a.useInfNorm; w.projectEveryStep = a.projectEveryStep;
        RkmStepResult r; r.errNorm = 0; r.projected = 0;
        int nproj = 0; double t = a.tcur[inst];
        for (int s = 0; s < a.nsteps; ++s) { r = tpiRkmStep<true>(c, w, a.h, cy); nproj += r.projected; t += a.h; }
        a.tcur[inst] = t;
        a.errNorm[inst] = r.errNorm;
        a.projCount[inst] += nproj;
        if (c.status && !(r.errNorm == r.errNorm)) *c.status |= 1;   // NaN error norm
    }
}

// Register-resident fused plan: the whole multi-step RKM loop of one instance in one thread.
template <class E>
__global__ void __launch_bounds__(128) fusedRkmKernel(const KArgs a) {
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
    for (int s = 0; s < a.nsteps; ++s) err = fusedRkmStep(e, y, a.h, a.useInfNorm);
#pragma unroll
    for (int i = 0; i < E::NY; ++i) a.y[(long long)i*a.N + inst] = y[i];
    a.tcur[inst] += a.nsteps*a.h;
    a.errNorm[inst] = err;
    if (a.status && !(err == err)) a.status[inst] |= 1;
}
template <class E>
cudaError_t launchFusedT(const KArgs& a, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(fusedRkmKernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.tableBytes);
    if (e != cudaSuccess) return e;
    fusedRkmKernel<E><<<(a.N + 127)/128, 128, a.tableBytes, stream>>>(a);
    return cudaGetLastError();
}

template <int OP>
cudaError_t launchOp(const KArgs& a, cudaStream_t stream) {
    const int grid = (a.N + TPI_THREADS - 1)/TPI_THREADS;
    if (a.stageInSmem) {
        cudaError_t e = cudaFuncSetAttribute(tpiKernel<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.tableBytes);
        if (e != cudaSuccess) return e;
        tpiKernel<OP, true><<<grid, TPI_THREADS, a.tableBytes, stream>>>(a);
    } else {
        tpiKernel<OP, false><<<grid, TPI_THREADS, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

__global__ void initGroundKernel(double* cache, int N) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= N) return;
    for (int f = 0; f < F_H; ++f) cache[(long long)f*N + k] = (f == F_XGB || f == F_XGB+4 || f == F_XGB+8) ? 1.0 : 0.0;
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
        const double* rec = a.cache + bodies[b].cacheBase + k;
        for (int i = 0; i < width; ++i) out[((long long)b*width + i)*a.N + k] = rec[(long long)(fieldOffset + i)*a.N];
    }
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
        case OP_KIN:     return launchOp<OP_KIN>(a, stream);
        case OP_ABI:     return launchOp<OP_ABI>(a, stream);
        case OP_EVAL:    return launchOp<OP_EVAL>(a, stream);
        case OP_CALCACC: return launchOp<OP_CALCACC>(a, stream);
        case OP_MULM:    return launchOp<OP_MULM>(a, stream);
        case OP_MULMINV: return launchOp<OP_MULMINV>(a, stream);
        case OP_RESID:   return launchOp<OP_RESID>(a, stream);
        case OP_RKM:     return launchOp<OP_RKM>(a, stream);
    }
    return cudaErrorInvalidValue;
}
bool fusedPlanSupports(int nb, const int* joints) {
    auto simple = [](int j) { return j == JT_PIN || j == JT_SLIDER; };
    if (nb == 2) return simple(joints[1]) || joints[1] == JT_UNIVERSAL;
    if (nb == 3) return simple(joints[1]) && simple(joints[2]);
    return false;
}
cudaError_t launchFusedRkm(const KArgs& a, const int* joints, cudaStream_t stream) {
    if (a.nb == 2) {
        switch (joints[1]) {
            case JT_PIN:       return launchFusedT<Chain1<JT_PIN>>(a, stream);
            case JT_SLIDER:    return launchFusedT<Chain1<JT_SLIDER>>(a, stream);
            case JT_UNIVERSAL: return launchFusedT<Chain1<JT_UNIVERSAL>>(a, stream);
        }
    } else if (a.nb == 3) {
        const int k = joints[1]*10 + joints[2];
        switch (k) {
            case JT_PIN*10 + JT_PIN:       return launchFusedT<Chain2<JT_PIN, JT_PIN>>(a, stream);
            case JT_PIN*10 + JT_SLIDER:    return launchFusedT<Chain2<JT_PIN, JT_SLIDER>>(a, stream);
            case JT_SLIDER*10 + JT_PIN:    return launchFusedT<Chain2<JT_SLIDER, JT_PIN>>(a, stream);
            case JT_SLIDER*10 + JT_SLIDER: return launchFusedT<Chain2<JT_SLIDER, JT_SLIDER>>(a, stream);
        }
    }
    return cudaErrorInvalidValue;
}
cudaError_t launchInitGround(double* cache, int N, cudaStream_t stream) {
    initGroundKernel<<<(N + 255)/256, 256, 0, stream>>>(cache, N);
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
cudaError_t launchDfmaProbe(double* out, int iters, int blocks, int threads, cudaStream_t stream) {
    dfmaProbeKernel<<<blocks, threads, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

} // namespace sbkd
