// sbk_rkm.cuh -- Runge-Kutta-Merson step for one instance (thread-per-instance plan).
//
// Replaces, for fixed step size (Integrator::setFixedStepSize + setAllowInterpolation(false)):
//   RungeKuttaMersonIntegratorRep::attemptODEStep   SimTKmath/Integrators/src/RungeKuttaMersonIntegrator.cpp:86-140
//   AbstractIntegratorRep::attemptDAEStep           AbstractIntegratorRep.cpp:137-208
//   IntegratorRep::calcErrorNorm / scaleDQ / calcRelativeScaling   IntegratorRep.h:454-513,644-655
//   SimbodyMatterSubsystemRep::projectQ (quaternion part) / normalizeQuaternions
//                                                   SimbodyMatterSubsystemRep.cpp:4096-4184,4448-4476
//   RBNodeBall/Free::enforceQuaternionConstraints   RigidBodyNodeSpec_Ball.h:417-434
// Each step costs exactly 5 derivative evaluations (f0 at the start of the step + 4 stages), as
// measured for the reference.
#pragma once
#include "sbk_sweeps.cuh"

namespace sbkd {

struct RkmWork {          // SoA work vectors, each [ny][N] with ny = nq + nu (q slots first)
    double* y;            // the resident state (q then u); == Ctx::q, Ctx::u = y + nq*sStride
    double* y0; double* f0; double* fa; double* fb; double* ys;
    double  accuracy, consTol;
    int     useInfNorm, projectEveryStep;
};

struct RkmStepResult { double errNorm; int projected; };

// Error norm of IntegratorRep::calcErrorNorm with UWeights = 1, no z.
// err lives in w.ys (overwritten by the caller with the error estimate), q1 = current state.
SBK_HD double rkmErrorNorm(const Ctx& c, const int inst, const RkmWork& w) {
    const int nq = c.nq, nu = c.nu;
    double qAcc = 0, uAcc = 0;
    // u part: uScale_i = |u0_i| > 1 ? 1/|u0_i| : 1   (calcRelativeScaling, frozen at step start)
#pragma unroll 8
    for (int i = 0; i < nu; ++i) {
        const double u0 = fabs(ldS(c, inst, w.y0, nq + i));
        const double sc = (u0*1.0 > 1.0) ? 1.0/u0 : 1.0;
        const double v  = sc*ldS(c, inst, w.ys, nq + i);
        if (w.useInfNorm) uAcc = fmax(uAcc, fabs(v)); else uAcc += v*v;
    }
    // q part: dqw = N * Wu * pinv(N) * dq (scaleDQ); identity except on quaternion slots
    for (int b = 1; b < c.nb; ++b) {
        const BodyConst& bc = c.bodies[b];
        int first = 0;
        if (bc.joint == JT_BALL || bc.joint == JT_FREE) {
            double q[4], e[4], o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = ldS(c, inst, w.y, bc.q0 + i); e[i] = ldS(c, inst, w.ys, bc.q0 + i); }
            const V3 du = quatNInvTimes(q, e);
            quatNTimes(q, du, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) { if (w.useInfNorm) qAcc = fmax(qAcc, fabs(o[i])); else qAcc += o[i]*o[i]; }
            first = 4;
        }
        const int nqb = bc.joint == JT_FREE ? 7 : bc.joint == JT_BALL ? 4 : bc.joint == JT_UNIVERSAL ? 2 : 1;
        for (int i = first; i < nqb; ++i) {
            const double v = ldS(c, inst, w.ys, bc.q0 + i);
            if (w.useInfNorm) qAcc = fmax(qAcc, fabs(v)); else qAcc += v*v;
        }
    }
    const double qNorm = w.useInfNorm ? qAcc : (nq ? sqrt(qAcc/nq) : 0.0);
    const double uNorm = w.useInfNorm ? uAcc : (nu ? sqrt(uAcc/nu) : 0.0);
    return qNorm >= uNorm ? qNorm : uNorm;
}

// The caller's Ctx must have q = w.y, u = w.y + nq*sStride and null qerr / fmobOut / FbodyOut.
template <bool LEAN>
SBK_HD RkmStepResult tpiRkmStep(const Ctx& c, const int inst, const RkmWork& w, const double h, double* cy) {
    const int nq = c.nq, ny = c.nq + c.nu;
    const long long uoff = (long long)nq*c.sStride;

    // f0 = f(y0): AbstractIntegratorRep.cpp:393 realizeStateDerivatives at the start of the step
    tpiEvalDerivatives<LEAN>(c, inst, cy, w.f0, w.f0 + uoff, nullptr);
#pragma unroll 8
    for (int i = 0; i < ny; ++i) {
        const double y0 = ldS(c, inst, w.y, i);
        stS(c, inst, w.y0, i, y0);
        stS(c, inst, w.y, i, y0 + (h/3)*ldS(c, inst, w.f0, i));
    }
    tpiEvalDerivatives<LEAN>(c, inst, cy, w.fa, w.fa + uoff, nullptr);                        // f1
#pragma unroll 8
    for (int i = 0; i < ny; ++i)
        stS(c, inst, w.y, i, ldS(c, inst, w.y0, i) + (h/6)*(ldS(c, inst, w.f0, i) + ldS(c, inst, w.fa, i)));
    tpiEvalDerivatives<LEAN>(c, inst, cy, w.fa, w.fa + uoff, nullptr);                        // f2 -> fa
#pragma unroll 8
    for (int i = 0; i < ny; ++i)
        stS(c, inst, w.y, i, ldS(c, inst, w.y0, i) + (h/8)*(ldS(c, inst, w.f0, i) + 3*ldS(c, inst, w.fa, i)));
    tpiEvalDerivatives<LEAN>(c, inst, cy, w.fb, w.fb + uoff, nullptr);                        // f3 -> fb
#pragma unroll 8
    for (int i = 0; i < ny; ++i) {
        const double ys = ldS(c, inst, w.y0, i) + (h/2)*(ldS(c, inst, w.f0, i) - 3*ldS(c, inst, w.fa, i) + 4*ldS(c, inst, w.fb, i));
        stS(c, inst, w.ys, i, ys); stS(c, inst, w.y, i, ys);
    }
    tpiEvalDerivatives<LEAN>(c, inst, cy, w.fa, w.fa + uoff, nullptr);                        // f4 -> fa
#pragma unroll 8
    for (int i = 0; i < ny; ++i) {
        const double y1 = ldS(c, inst, w.y0, i) + (h/6)*(ldS(c, inst, w.f0, i) + 4*ldS(c, inst, w.fb, i) + ldS(c, inst, w.fa, i));
        stS(c, inst, w.y, i, y1);
        stS(c, inst, w.ys, i, 0.2*fabs(y1 - ldS(c, inst, w.ys, i)));                            // y1err
    }

    RkmStepResult res; res.projected = 0;
    res.errNorm = rkmErrorNorm(c, inst, w);
    // attemptDAEStep: project only if errNorm <= 2^4 * accuracy (AbstractIntegratorRep.cpp:165-166)
    if (c.nquat > 0 && !(res.errNorm > 16.0*w.accuracy)) {
        double acc = 0;
        for (int b = 1; b < c.nb; ++b) {
            const BodyConst& bc = c.bodies[b];
            if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
            double n2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double qi = ldS(c, inst, w.y, bc.q0 + i); n2 += qi*qi; }
            const double e = sqrt(n2) - 1.0;
            if (w.useInfNorm) acc = fmax(acc, fabs(e)); else acc += e*e;
        }
        const double quatNorm = w.useInfNorm ? acc : sqrt(acc/c.nquat);
        if (quatNorm > w.consTol || w.projectEveryStep) {
            for (int b = 1; b < c.nb; ++b) {
                const BodyConst& bc = c.bodies[b];
                if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                double q[4], e[4], n2 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = ldS(c, inst, w.y, bc.q0 + i); e[i] = ldS(c, inst, w.ys, bc.q0 + i); n2 += q[i]*q[i]; }
                const double n = sqrt(n2);
                double dt = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = q[i]/n; dt += e[i]*q[i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { stS(c, inst, w.y, bc.q0 + i, q[i]); stS(c, inst, w.ys, bc.q0 + i, e[i] - dt*q[i]); }
            }
            res.projected = 1;
            res.errNorm = rkmErrorNorm(c, inst, w);      // takeOneStep recomputes it (AbstractIntegratorRep.cpp:556)
        }
    }
    return res;
}

} // namespace sbkd
