// sbk_fused.cuh -- register-resident fused plan for small serial chains (1 or 2 mobilizers).
//
// The whole Runge-Kutta-Merson multi-step loop of one instance runs in ONE thread with the
// state, the five RKM work vectors and every per-body quantity (X_GB, H, P, D^-1, G, z, ...)
// held in registers: no per-body cache in HBM at all.  Per launch the only DRAM traffic is the
// compulsory read and write of y (8*ny bytes each way per instance, amortised over nsteps).
// This is the plan behind BASELINE.json config 2 (2-link Pin pendulum, 1M instances).
//
// It calls exactly the same cores (kinCore / abiCore / zCore / accCore) as the cache-based
// plans, in the same order, so results are bitwise identical to the thread-per-instance plan.
#pragma once
#include "sbk_rkm.cuh"

namespace sbkd {

// Derivative evaluation for a 2-body chain  Ground -> b0 (J0) -> b1 (J1).
// y = [q(b0), q(b1), u(b0), u(b1)], ydot likewise.  Quaternion mobilizers are not handled here.
template <int J0, int J1>
struct Chain2 {
    enum { NQ0 = JointDims<J0>::nq, NQ1 = JointDims<J1>::nq, D0 = JointDims<J0>::nu, D1 = JointDims<J1>::nu,
           NQ = NQ0 + NQ1, NU = D0 + D1, NY = NQ + NU, NB = 3 };
    const BodyConst* b0; const BodyConst* b1; const ForceConst* forces; double gx, gy, gz;

    SBK_HD void eval(const double* y, double* ydot) const {
        const double* q0 = y; const double* q1 = y + NQ0; const double* u0 = y + NQ; const double* u1 = y + NQ + D0;
        double qe;
        KinOut<D0> k0; KinOut<D1> k1;
        SV V0 = zeroSV();
        kinCore<J0>(*b0, q0, u0, identity3(), zero3(), V0, k0, ydot, qe);
        kinCore<J1>(*b1, q1, u1, k0.R, k0.p, k0.V, k1, ydot + NQ0, qe);
        // inward: body 1 (leaf), then body 0
        AbiOut<D1> a1; AbiOut<D0> a0;
        double f1[D1], f0[D0], eps1[D1], eps0[D0]; SV zP1, zP0;
        abiCore<D1>(abiFromRigid(b1->mass, k1.c, k1.G), k1.H, k1.acor, k1.gyro, a1);
        mobilityForces<D1>(*b1, forces, q1, u1, f1);
        zCore<D1>(k1.H, a1.G, a1.zb - gravityForce(b1->mass, k1.c, gx, gy, gz), f1, eps1, zP1);
        ABI P0 = abiFromRigid(b0->mass, k0.c, k0.G);
        addInto(P0, shiftABI(a1.PP, k1.l));
        abiCore<D0>(P0, k0.H, k0.acor, k0.gyro, a0);
        mobilityForces<D0>(*b0, forces, q0, u0, f0);
        zCore<D0>(k0.H, a0.G, (a0.zb - gravityForce(b0->mass, k0.c, gx, gy, gz)) + phi(k1.l, zP1), f0, eps0, zP0);
        // outward
        SV A0, A1;
        accCore<D0, true>(k0.H, a0.G, a0.DI, eps0, k0.l, zeroSV(), k0.acor, ydot + NQ, A0);
        accCore<D1, true>(k1.H, a1.G, a1.DI, eps1, k1.l, A0, k1.acor, ydot + NQ + D0, A1);
    }
};

template <int J0>
struct Chain1 {
    enum { NQ = JointDims<J0>::nq, NU = JointDims<J0>::nu, NY = NQ + NU, NB = 2 };
    const BodyConst* b0; const BodyConst* b1; const ForceConst* forces; double gx, gy, gz;

    SBK_HD void eval(const double* y, double* ydot) const {
        const double* q0 = y; const double* u0 = y + NQ;
        double qe; KinOut<NU> k0;
        kinCore<J0>(*b0, q0, u0, identity3(), zero3(), zeroSV(), k0, ydot, qe);
        AbiOut<NU> a0; double f0[NU], eps0[NU]; SV zP0, A0;
        abiCore<NU>(abiFromRigid(b0->mass, k0.c, k0.G), k0.H, k0.acor, k0.gyro, a0);
        mobilityForces<NU>(*b0, forces, q0, u0, f0);
        zCore<NU>(k0.H, a0.G, a0.zb - gravityForce(b0->mass, k0.c, gx, gy, gz), f0, eps0, zP0);
        accCore<NU, true>(k0.H, a0.G, a0.DI, eps0, k0.l, zeroSV(), k0.acor, ydot + NQ, A0);
    }
};

// One RKM attempt entirely in registers (same arithmetic as tpiRkmStep; no quaternions here so the
// q-part of the error norm is the plain RMS and there is no projection).  The five derivative
// evaluations share ONE inlined copy of eval() inside a loop over the stages: no call boundary (and
// no generic loads through a `this` pointer) in the step, one copy of the sweeps in the instruction
// cache.  fresh = true: start of a step, evaluates f0 = f(y) and saves y0 = y; fresh = false: retry
// of a failed attempt from the saved y0 / f0 with a new h.
template <class E>
SBK_HD double fusedRkmAttempt(const E& e, double* y0, double* f0, double* y, const double h, const int useInfNorm, const bool fresh) {
    constexpr int NY = E::NY, NQ = E::NQ, NU = E::NU;
    double fa[NY], fb[NY], ys[NY], ft[NY];
    double qAcc = 0, uAcc = 0;
    if (!fresh) {
#pragma unroll
        for (int i = 0; i < NY; ++i) y[i] = y0[i] + (h/3)*f0[i];
    }
#pragma unroll 1
    for (int stage = fresh ? 0 : 1; stage < 5; ++stage) {
        e.eval(y, ft);
        if (stage == 0) {
#pragma unroll
            for (int i = 0; i < NY; ++i) { f0[i] = ft[i]; y0[i] = y[i]; y[i] = y0[i] + (h/3)*f0[i]; }
        } else if (stage == 1) {
#pragma unroll
            for (int i = 0; i < NY; ++i) { fa[i] = ft[i]; y[i] = y0[i] + (h/6)*(f0[i] + fa[i]); }
        } else if (stage == 2) {
#pragma unroll
            for (int i = 0; i < NY; ++i) { fa[i] = ft[i]; y[i] = y0[i] + (h/8)*(f0[i] + 3*fa[i]); }
        } else if (stage == 3) {
#pragma unroll
            for (int i = 0; i < NY; ++i) { fb[i] = ft[i]; ys[i] = y0[i] + (h/2)*(f0[i] - 3*fa[i] + 4*fb[i]); y[i] = ys[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < NY; ++i) {
                const double y1 = y0[i] + (h/6)*(f0[i] + 4*fb[i] + ft[i]);
                y[i] = y1;
                const double err = 0.2*fabs(y1 - ys[i]);
                if (i < NQ) qAcc = normAcc(qAcc, err, useInfNorm);
                else {
                    const double a0 = fabs(y0[i]);
                    const double sc = (a0*1.0 > 1.0) ? 1.0/a0 : 1.0;
                    const double v = sc*err;
                    uAcc = normAcc(uAcc, v, useInfNorm);
                }
            }
        }
    }
    const double qNorm = useInfNorm ? qAcc : sqrt(qAcc/NQ);
    const double uNorm = useInfNorm ? uAcc : sqrt(uAcc/NU);
    return normMax(uNorm, qNorm);
}
// One fixed-size step.
template <class E>
SBK_HD double fusedRkmStep(const E& e, double* y, const double h, const int useInfNorm) {
    constexpr int NY = E::NY;
    double y0[NY], f0[NY];
    return fusedRkmAttempt(e, y0, f0, y, h, useInfNorm, true);
}
// Error-controlled stepping to tFinal (cf. tpiRkmAdaptive).
template <class E>
SBK_HD void fusedRkmAdaptive(const E& e, double* y, const StepLimits& lim, const double tFinal, const int allowInterpolation,
                             const int maxAttempts, const int useInfNorm, AdaptiveState& st, double& lastErr) {
    constexpr int NY = E::NY;
    int budget = maxAttempts; bool fresh = true;
    double y0[NY], f0[NY];
    const WarpVote vote;
    for (;;) {                                   // one attempt per trip; see WarpVote (sbk_rkm.cuh)
        const bool mine = st.t < tFinal && budget > 0;
        if (!vote(mine)) break;
        if (mine) {
            bool limited = false; double t1;
            if (allowInterpolation) t1 = st.t + st.h;
            else if (tFinal < st.t + 0.95*st.h)  { limited = true; t1 = tFinal; }
            else if (tFinal > st.t + 1.001*st.h) t1 = st.t + st.h;
            else t1 = tFinal;
            lastErr = fusedRkmAttempt(e, y0, f0, y, t1 - st.t, useInfNorm, fresh);
            ++st.attempts; --budget;
            fresh = adjustStepSize(lastErr, lim, limited, st.h);
            if (fresh) { st.lastStep = t1 - st.t; st.t = t1; ++st.steps; }
        }
    }
    if (!fresh) {
#pragma unroll
        for (int i = 0; i < NY; ++i) y[i] = y0[i];
    }
}

} // namespace sbkd
