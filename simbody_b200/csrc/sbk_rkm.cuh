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
#include "sbk_local.cuh"

namespace sbkd {

struct RkmWork {          // SoA work vectors, each [ny][N] with ny = nq + nu (q slots first)
    double* y;            // the resident state (q then u); == Ctx::q, Ctx::u = y + nq*sStride
    double* y0; double* f0; double* fa; double* fb; double* ys;
    double  accuracy, consTol;
    int     useInfNorm, projectEveryStep;
};

struct RkmStepResult { double errNorm; int projected; };

// Error-norm accumulation (IntegratorRep.h:454-488; weightedNormRMS / normInf): sum of squares, or the largest magnitude in
// Inf-norm mode.  A NaN must survive the Inf-norm (fmax would drop it): adjustStepSize's isFinite test (AbstractIntegratorRep.cpp:
// 451-456) has to see a non-finite norm for a diverged instance.
SBK_HD double normAcc(const double acc, const double v, const int useInf) {
    const double t = useInf ? fabs(v) : v*v;
    return useInf ? ((t > acc || t != t) ? t : acc) : acc + t;
}
SBK_HD double normMax(const double a, const double b) { return (b > a || b != b) ? b : a; }     // NaN-propagating max (a NaN a stays)
SBK_HD bool finiteNorm(const double e) { return fabs(e) <= 1.7976931348623157e308; }
// attemptDAEStep's projection limit (AbstractIntegratorRep.cpp:165-190): a constraint violation beyond max(2 tol, sqrt(tol)) is a
// convergence failure of the step (error norm = Infinity, no projection)
SBK_HD double projectionLimit(const double consTol) { const double r = sqrt(consTol); return 2*consTol > r ? 2*consTol : r; }

// Error norm of IntegratorRep::calcErrorNorm with UWeights = 1, no z.
// err lives in w.ys (overwritten by the caller with the error estimate), q1 = current state.
template <bool BLK, class TBL>
SBK_HD double rkmErrorNorm(const Ctx& c, const TBL& T, const int inst, const RkmWork& w) {
    const int nq = c.nq, nu = c.nu;
    double qAcc = 0, uAcc = 0;
    // u part: uScale_i = |u0_i| > 1 ? 1/|u0_i| : 1   (calcRelativeScaling, frozen at step start)
#pragma unroll 8
    for (int i = 0; i < nu; ++i) {
        const double u0 = fabs(ldS<BLK>(c, inst, w.y0, nq + i));
        const double sc = (u0*1.0 > 1.0) ? 1.0/u0 : 1.0;
        const double v  = sc*ldS<BLK>(c, inst, w.ys, nq + i);
        uAcc = normAcc(uAcc, v, w.useInfNorm);
    }
    // q part: dqw = N * Wu * pinv(N) * dq (scaleDQ); identity except on quaternion slots
    for (int b = 1; b < c.nb; ++b) {
        const auto& bc = T.bodies[b];
        int first = 0;
        if (bc.joint == JT_BALL || bc.joint == JT_FREE) {
            double q[4], e[4], o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = ldS<BLK>(c, inst, w.y, bc.q0 + i); e[i] = ldS<BLK>(c, inst, w.ys, bc.q0 + i); }
            const V3 du = quatNInvTimes(q, e);
            quatNTimes(q, du, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) qAcc = normAcc(qAcc, o[i], w.useInfNorm);
            first = 4;
        }
        const int nqb = nqOfJoint(bc.joint);
        for (int i = first; i < nqb; ++i) {
            qAcc = normAcc(qAcc, ldS<BLK>(c, inst, w.ys, bc.q0 + i), w.useInfNorm);
        }
    }
    const double qNorm = w.useInfNorm ? qAcc : (nq ? sqrt(qAcc/nq) : 0.0);
    const double uNorm = w.useInfNorm ? uAcc : (nu ? sqrt(uAcc/nu) : 0.0);
    return normMax(uNorm, qNorm);
}

// AbstractIntegratorRep::adjustStepSize (AbstractIntegratorRep.cpp:448-502).  `h` is the current
// step size (in/out); returns true ("success") iff the new step is at least as big as the old one.
struct StepLimits { double accuracy, minStep, maxStep; };   // minStep/maxStep <= 0: no user limit
SBK_HD bool adjustStepSize(const double err, const StepLimits& lim, const bool hWasArtificiallyLimited, double& h) {
    const double Safety = 0.9, MinShrink = 0.1, MaxGrow = 5, HysteresisLow = 0.9, HysteresisHigh = 1.2;
    const double cur = h; double nw;
    if (!(fabs(err) <= 1.7976931348623157e308)) nw = MinShrink*cur;           // !isFinite(err)
    else if (err == 0) nw = MaxGrow*cur;
    else nw = Safety*cur*pow(lim.accuracy/err, 1.0/4.0);                      // errOrder = 4 for RKM
    if (nw > cur) { if (hWasArtificiallyLimited || nw < HysteresisHigh*cur) nw = cur; }
    if (nw < cur) { if (err <= lim.accuracy) nw = cur; else nw = fmin(nw, HysteresisLow*cur); }
    nw = fmin(nw, MaxGrow*cur); nw = fmax(nw, MinShrink*cur);
    if (lim.minStep > 0) nw = fmax(nw, lim.minStep);
    if (lim.maxStep > 0) nw = fmin(nw, lim.maxStep);
    h = nw;
    return nw >= cur;
}

// out(i, v) for every slot i with v[j] = src_j[i]; operands are loaded a chunk of slots at a time.
template <bool BLK, int NIN, class Fn>
SBK_HD void rkmCombine(const Ctx& c, const int inst, const int ny, const double* s0, const double* s1, const double* s2,
                       const double* s3, const double* s4, Fn out) {
    constexpr int CH = NIN >= 4 ? 4 : 8;
    const double* src[5] = {s0, s1, s2, s3, s4};
    for (int i0 = 0; i0 < ny; i0 += CH) {
        double v[CH][NIN];
#pragma unroll
        for (int k = 0; k < CH; ++k)
#pragma unroll
            for (int j = 0; j < NIN; ++j) v[k][j] = (i0 + k < ny) ? ldS<BLK>(c, inst, src[j], i0 + k) : 0.0;
#pragma unroll
        for (int k = 0; k < CH; ++k) if (i0 + k < ny) out(i0 + k, v[k]);
    }
}

// One RKM attempt.  FRESH = true: start of a step -- evaluate f0 = f(y) and save y0 = y
// (AbstractIntegratorRep.cpp:390-396).  FRESH = false: retry of a failed attempt with a smaller h
// from the saved y0 / f0 (no re-evaluation, as in takeOneStep's do/while).
// The caller's Ctx must have q = w.y, u = w.y + nq*sStride and null qerr / fmobOut / FbodyOut.
// The derivative evaluation behind a table type: Tables = ground-frame sweeps (FULL records or LEAN reversible
// kinematics), LTables = body-frame sweeps (sbk_local.cuh).
template <bool LEAN, int JMASK, bool LOCK = false> SBK_HD void rkmEval(const Ctx& c, const Tables& T, const int inst, double* cy, double* qd, double* ud, const bool on = true) {
    tpiEvalDerivatives<LEAN, JMASK, false, LOCK>(c, T, inst, cy, qd, ud, nullptr, on);
}
template <bool LEAN, int JMASK, bool LOCK = false> SBK_HD void rkmEval(const Ctx& c, const LTables& T, const int inst, double* cy, double* qd, double* ud, const bool on = true) {
    static_assert(!LOCK, "the body-frame integrator has its own lockstep attempt (sbk_lrkm.cuh)");
    if (on) lEvalDerivatives<JMASK>(c, T, inst, cy, qd, ud);
}

// LOCK (CTA-voting error-controlled kernel): every thread of the CTA walks every evaluation some thread needs -- `anyFresh` = some
// thread starts a step and needs f0 -- and meets the others at every body; mine = false: nothing to do but keep pace.
template <bool LEAN, int JMASK = JM_ALL, class TBL = Tables, bool LOCK = false>
SBK_HD RkmStepResult tpiRkmStep(const Ctx& c, const TBL& T, const int inst, const RkmWork& w, const double h, double* cy, const bool fresh = true,
                                const bool mine = true, const bool anyFresh = true) {
    constexpr bool BLK = LEAN && SBK_DEV_BLK;
    const int nq = c.nq, ny = c.nq + c.nu;
    const long long uoff = BLK ? (long long)nq*BLK_LANES : (long long)nq*c.sStride;   // u rows follow the q rows

    // The five derivative evaluations share ONE inlined call site (a loop over the stages): no
    // function-call boundary with its register save/restore inside the step, and one copy of the
    // sweeps in the instruction cache.  Stage combinations: all operands of a chunk of slots are
    // requested before the first store, so a chunk costs one trip to memory (the arrays may alias
    // as far as the compiler knows, and load-store-load-store would serialise on memory latency).
#pragma unroll 1
    for (int stage = 0; stage < 5; ++stage) {
        double* fdst = stage == 0 ? w.f0 : (stage == 3 ? w.fb : w.fa);
        // stage 0: f0 = f(y0), AbstractIntegratorRep.cpp:393 realizeStateDerivatives at the start of a step;
        // a retry after a failed attempt keeps the saved y0 / f0
        if constexpr (LOCK) { if (stage > 0 || anyFresh) rkmEval<LEAN, JMASK, true>(c, T, inst, cy, fdst, fdst + uoff, mine && (stage > 0 || fresh)); }
        else if (stage > 0 || fresh) rkmEval<LEAN, JMASK>(c, T, inst, cy, fdst, fdst + uoff);
        if (!mine) continue;
        if (stage == 0) {
            if (fresh) rkmCombine<BLK, 2>(c, inst, ny, w.y, w.f0, nullptr, nullptr, nullptr, [&](int i, const double* v) {
                           stS<BLK>(c, inst, w.y0, i, v[0]);
                           stS<BLK>(c, inst, w.y, i, v[0] + (h/3)*v[1]); });
            else       rkmCombine<BLK, 2>(c, inst, ny, w.y0, w.f0, nullptr, nullptr, nullptr, [&](int i, const double* v) {
                           stS<BLK>(c, inst, w.y, i, v[0] + (h/3)*v[1]); });
        } else if (stage == 1) {                                                             // fa = f1
            rkmCombine<BLK, 3>(c, inst, ny, w.y0, w.f0, w.fa, nullptr, nullptr, [&](int i, const double* v) {
                stS<BLK>(c, inst, w.y, i, v[0] + (h/6)*(v[1] + v[2])); });
        } else if (stage == 2) {                                                             // fa = f2
            rkmCombine<BLK, 3>(c, inst, ny, w.y0, w.f0, w.fa, nullptr, nullptr, [&](int i, const double* v) {
                stS<BLK>(c, inst, w.y, i, v[0] + (h/8)*(v[1] + 3*v[2])); });
        } else if (stage == 3) {                                                             // fb = f3
            rkmCombine<BLK, 4>(c, inst, ny, w.y0, w.f0, w.fa, w.fb, nullptr, [&](int i, const double* v) {
                const double ys = v[0] + (h/2)*(v[1] - 3*v[2] + 4*v[3]);
                stS<BLK>(c, inst, w.ys, i, ys); stS<BLK>(c, inst, w.y, i, ys); });
        } else {                                                                             // fa = f4
            rkmCombine<BLK, 5>(c, inst, ny, w.y0, w.f0, w.fb, w.fa, w.ys, [&](int i, const double* v) {
                const double y1 = v[0] + (h/6)*(v[1] + 4*v[2] + v[3]);
                stS<BLK>(c, inst, w.y, i, y1);
                stS<BLK>(c, inst, w.ys, i, 0.2*fabs(y1 - v[4])); });                        // y1err
        }
    }

    RkmStepResult res; res.projected = 0; res.errNorm = 0;
    if (!mine) return res;
    res.errNorm = rkmErrorNorm<BLK>(c, T, inst, w);
    // attemptDAEStep: project only if errNorm <= 2^4 * accuracy (AbstractIntegratorRep.cpp:165-166)
    if (c.nquat > 0 && !(res.errNorm > 16.0*w.accuracy)) {
        double acc = 0;
        for (int b = 1; b < c.nb; ++b) {
            const auto& bc = T.bodies[b];
            if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
            double n2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double qi = ldS<BLK>(c, inst, w.y, bc.q0 + i); n2 += qi*qi; }
            const double e = sqrt(n2) - 1.0;
            acc = normAcc(acc, e, w.useInfNorm);
        }
        const double quatNorm = w.useInfNorm ? acc : sqrt(acc/c.nquat);
        if (quatNorm > projectionLimit(w.consTol)) res.errNorm = 1.0/0.0;      // convergence failure: too far off the manifold to project
        else if (quatNorm > w.consTol || w.projectEveryStep) {
            for (int b = 1; b < c.nb; ++b) {
                const auto& bc = T.bodies[b];
                if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                double q[4], e[4], n2 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = ldS<BLK>(c, inst, w.y, bc.q0 + i); e[i] = ldS<BLK>(c, inst, w.ys, bc.q0 + i); n2 += q[i]*q[i]; }
                const double n = sqrt(n2);
                double dt = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = q[i]/n; dt += e[i]*q[i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { stS<BLK>(c, inst, w.y, bc.q0 + i, q[i]); stS<BLK>(c, inst, w.ys, bc.q0 + i, e[i] - dt*q[i]); }
            }
            res.projected = 1;
            res.errNorm = rkmErrorNorm<BLK>(c, T, inst, w);      // takeOneStep recomputes it (AbstractIntegratorRep.cpp:556)
        }
    }
    return res;
}

// Integrator::stepTo(tFinal) with error control for ONE instance (AbstractIntegratorRep.cpp:216-368,
// 513-578): internal steps until the advanced time reaches tFinal.  With allowInterpolation (the
// reference's default) steps are never shortened to hit tFinal, so the advanced state ends at
// t >= tFinal; without it the last step lands on tFinal (hWasArtificiallyLimited logic).
struct AdaptiveState { double t, h, lastStep; int steps, attempts; };
// The error-controlled loops make ONE attempt per trip and every lane of the warp keeps tripping (predicated off once its
// instance is done) until no lane has work left: the lanes reconverge at each attempt (WarpVote below).  With the reference's nested
// `while (t < tFinal) do attempt while (!ok)` form a lane that had rejected a step never rejoined the others before the end of
// the outer loop -- the warp then ran its lanes' attempts one group after the other (measured: 25x slower on the humanoid).
// vote(mine) -> "some thread of the group still has work": the group is the warp (WarpVote) or the whole CTA (CtaVote, every
// thread of the CTA must call it).  The CTA form also keeps the CTA's warps in the same region of a 300 KB instruction stream:
// with the warps of an SM spread over it, 80% of the stall samples of the humanoid's kernel were instruction-fetch misses (ncu).
SBK_HD void ctaBarrier() {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
}
struct WarpVote {
    static constexpr bool CTA = false;
    unsigned lanes;
    SBK_HD WarpVote() : lanes(1u) {
#ifdef __CUDA_ARCH__
        lanes = __activemask();
#endif
    }
    SBK_HD bool operator()(const bool mine) const {
#ifdef __CUDA_ARCH__
        return __any_sync(lanes, mine) != 0;
#else
        return mine;
#endif
    }
};
struct CtaVote {
    static constexpr bool CTA = true;
    SBK_HD bool operator()(const bool mine) const {
#ifdef __CUDA_ARCH__
        return __syncthreads_or(mine ? 1 : 0) != 0;
#else
        return mine;
#endif
    }
};
template <bool LEAN, int JMASK = JM_ALL, class TBL = Tables, class VOTE = WarpVote>
SBK_HD void tpiRkmAdaptive(const Ctx& c, const TBL& T, const int inst, const RkmWork& w, const StepLimits& lim, const double tFinal,
                           const int allowInterpolation, const int maxAttempts, AdaptiveState& st, double* cy,
                           double& lastErr, int& nproj, const bool live = true, const VOTE vote = VOTE()) {
    constexpr bool BLK = LEAN && SBK_DEV_BLK;
    int budget = maxAttempts; bool fresh = true;
    for (;;) {                                   // one attempt per trip; see WarpVote
        const bool mine = live && st.t < tFinal && budget > 0;
        if (!vote(mine)) break;
        bool limited = false; double t1 = st.t;
        if (mine) {
            if (allowInterpolation) t1 = st.t + st.h;
            else if (tFinal < st.t + 0.95*st.h)  { limited = true; t1 = tFinal; }
            else if (tFinal > st.t + 1.001*st.h) t1 = st.t + st.h;
            else t1 = tFinal;
        }
        const double hTry = t1 - st.t;
        RkmStepResult r; r.errNorm = 0; r.projected = 0;
        if constexpr (VOTE::CTA) { const bool anyFresh = vote(mine && fresh); r = tpiRkmStep<LEAN, JMASK, TBL, true>(c, T, inst, w, hTry, cy, fresh, mine, anyFresh); }
        else if (mine) r = tpiRkmStep<LEAN, JMASK>(c, T, inst, w, hTry, cy, fresh);
        if (mine) {
            ++st.attempts; --budget; nproj += r.projected; lastErr = r.errNorm;
            fresh = adjustStepSize(r.errNorm, lim, limited, st.h);
            if (fresh) { st.lastStep = t1 - st.t; st.t = t1; ++st.steps; }
        }
    }
    if (!fresh) {    // out of budget inside a failing step: put y0 back, report through the status word
#pragma unroll 8
        for (int i = 0; i < c.nq + c.nu; ++i) stS<BLK>(c, inst, w.y, i, ldS<BLK>(c, inst, w.y0, i));
    }
}

} // namespace sbkd
