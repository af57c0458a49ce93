// sbk_lrkm.cuh -- Runge-Kutta-Merson step on the body-frame sweeps, TWO sweeps per derivative evaluation.
//
// Same step as tpiRkmStep (RungeKuttaMersonIntegrator.cpp:86-140, AbstractIntegratorRep.cpp:137-208,
// IntegratorRep.h:454-513): 5 evaluations, the same stage combinations in the same operand order, the
// same error norm and quaternion projection rule.  What is fused: the outward (acceleration) sweep of
// evaluation k yields udot, qdot of a body while it still holds the body's q, u -- so the body step goes
// straight on to that body's slots of the NEXT stage state (y0 + h * sum c_i f_i needs only this body's
// slots of y0, f0, f2, f3), and from the new q, u to the sin/cos and the velocity recurrence of the next
// evaluation.  One evaluation is then an inward and an outward sweep; the stage vectors are touched once
// (f1 and f4 never reach memory), and the error norm accumulates in the last outward sweep.
//
// Arrays (CTA-blocked [block][slot][lane] on the device):
//   Y   accepted state y0 of the step          W   stage state (y0 + ...), finally the error estimate
//   F0, F2, F3  derivative vectors that a later stage needs     Ynext  where y1 goes (== Y for fixed steps)
#pragma once
#include "sbk_rkm.cuh"

namespace sbkd {

struct LRkmWork {
    double* Y; double* W; double* F0; double* F2; double* F3; double* Ynext;
    double accuracy, consTol; int useInfNorm, projectEveryStep;
};
// carry rows of the fused outward sweep: v of the evaluated state, a, v of the next stage's state, then the
// running error sums of the last stage (q part, u part, sum of (|quat| - 1)^2)
// then the two prefetch slots (sbk_local.cuh)
enum { LF_V = 0, LF_A = 6, LF_V2 = 12, LF_QACC = 18, LF_UACC = 19, LF_QUATACC = 20, LF_PF = 33, LFCARRY_ROWS = LF_PF + 2*LPF_ROWS };

// Stage combination as data (RungeKuttaMersonIntegrator.cpp:86-140): the next stage state is
//   y0 + hk * (((a0*f0 + m2*f2) + m3*f3) + mf*f)      f = this evaluation's derivative
// which evaluates to the reference's expressions term by term (absent terms contribute an exact 0).
struct LStage { double hk, a0, m2, m3, mf; double* fdst; int stage; };
SBK_HD LStage lstageOf(const int stage, const double h, const LRkmWork& w) {
    LStage s; s.stage = stage; s.a0 = 1; s.m2 = 0; s.m3 = 0; s.mf = 1; s.fdst = nullptr; s.hk = h/6;
    if (stage == 0)      { s.hk = h/3; s.a0 = 0; s.fdst = w.F0; }      // y0 + h/3 f0
    else if (stage == 1) { s.hk = h/6; }                               // y0 + h/6 (f0 + f1)
    else if (stage == 2) { s.hk = h/8; s.mf = 3; s.fdst = w.F2; }      // y0 + h/8 (f0 + 3 f2)
    else if (stage == 3) { s.hk = h/2; s.m2 = -3; s.mf = 4; s.fdst = w.F3; }   // ys = y0 + h/2 (f0 - 3 f2 + 4 f3)
    else if (stage == 4) { s.hk = h/6; s.m3 = 4; }                     // y1 = y0 + h/6 (f0 + 4 f3 + f4)
    return s;
}

// Request the rows the fused outward step of body bc will read (layout LPfOut<kind of bc>).
template <int JT, bool BLK>
SBK_HD void lPrefetchOutT(const Ctx& c, const LBody& bc, const int inst, double* pf, const LRkmWork& w, const double* S, const LStage& sg) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    typedef LPfOut<JT> L;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    const long long rs = BLK ? BLK_LANES : me.stride, ss = BLK ? BLK_LANES : c.sStride;
    const long long offQ = stateIndex<BLK>(c, inst, bc.q0), offU = stateIndex<BLK>(c, inst, c.nq + bc.u0);
    lpfRowsN<NQ>(pf + L::QU*SBK_CARRY_STRIDE, S + offQ, ss);
    lpfRowsN<d>(pf + (L::QU + NQ)*SBK_CARRY_STRIDE, S + offU, ss);
    if (sg.stage < 0) return;
    lpfRowsN<lgCount<JT>()>(pf + L::G*SBK_CARRY_STRIDE, me.p + LR_G*rs, rs);
    lpfRowsN<d>(pf + L::NU*SBK_CARRY_STRIDE, me.p + lrNU(d)*rs, rs);
    lpfRowsN<lscCount<JT>()>(pf + L::SC*SBK_CARRY_STRIDE, me.p + LR_SC*rs, rs);
    if (sg.stage > 0) {
        lpfRowsN<NQ>(pf + L::Y*SBK_CARRY_STRIDE, w.Y + offQ, ss);
        lpfRowsN<d>(pf + (L::Y + NQ)*SBK_CARRY_STRIDE, w.Y + offU, ss);
        if constexpr (L::HAS_F0 != 0) {
            lpfRowsN<NQ>(pf + L::F0*SBK_CARRY_STRIDE, w.F0 + offQ, ss);
            lpfRowsN<d>(pf + (L::F0 + NQ)*SBK_CARRY_STRIDE, w.F0 + offU, ss);
        }
        if constexpr (L::HAS_F23 != 0) {
            const double* f23 = sg.m2 != 0 ? w.F2 : sg.m3 != 0 ? w.F3 : nullptr;
            if (f23) {
                lpfRowsN<NQ>(pf + L::F23*SBK_CARRY_STRIDE, f23 + offQ, ss);
                lpfRowsN<d>(pf + (L::F23 + NQ)*SBK_CARRY_STRIDE, f23 + offU, ss);
            }
        }
    }
}
template <int JMASK, bool BLK>
SBK_HD void lPrefetchOut(const Ctx& c, const LBody& bc, const int inst, double* pf, const LRkmWork& w, const double* S, const LStage& sg) {
    SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lPrefetchOutT<JT, BLK>(c, bc, inst, pf, w, S, sg)));
}

// One body of the fused outward sweep.  Phase 0 (stage >= 0): acceleration of the evaluation at state S (S = Y for
// stage 0, W otherwise) and the body's slots of the next state; phase 1: sin/cos and velocity of that next state.
// Both phases run the same joint / velocity code (a two-trip loop, one copy in the instruction cache).  stage < 0:
// phase 1 only, on the state in S (the stand-alone velocity sweep).  vr / vw: velocity buffer read / written.
template <int JT>
SBK_BODY void lFusedOutBody(const Ctx& c, const LBody& bc, const int inst, double* cy, const LRkmWork& w, const double* S,
                            const LStage& sg, const int vr, const int vw, const double* pf) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    typedef LPfOut<JT> L;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    const int nq = c.nq, stage = sg.stage;
    double qu[dim1(NQ + d)], up[dim1(d)], sc[LSC_ROWS];
#pragma unroll
    for (int i = 0; i < NQ + d; ++i) qu[i] = pf[(L::QU + i)*SBK_CARRY_STRIDE];
    if (stage >= 0) {
#pragma unroll
        for (int i = 0; i < lscCount<JT>(); ++i) sc[i] = pf[(L::SC + i)*SBK_CARRY_STRIDE];
    }
    // stage vectors that were too many rows for the slot: requested now, consumed after the acceleration arithmetic
    double f0d[dim1(NQ + d)], f23d[dim1(NQ + d)];
    if (stage > 0) {
        if constexpr (L::HAS_F0 == 0) {
#pragma unroll
            for (int i = 0; i < NQ + d; ++i) f0d[i] = ldS<BLK>(c, inst, w.F0, i < NQ ? bc.q0 + i : nq + bc.u0 + (i - NQ));
        }
        if constexpr (L::HAS_F23 == 0) {
            const double* f23 = sg.m2 != 0 ? w.F2 : sg.m3 != 0 ? w.F3 : nullptr;
#pragma unroll
            for (int i = 0; i < NQ + d; ++i) f23d[i] = f23 ? ldS<BLK>(c, inst, f23, i < NQ ? bc.q0 + i : nq + bc.u0 + (i - NQ)) : 0.0;
        }
    }
#pragma unroll 1
    for (int ph = stage < 0 ? 1 : 0; ph < 2; ++ph) {
        if (ph) {
            if constexpr (JT == JT_PIN) sincos(qu[0], &sc[0], &sc[1]);
            if constexpr (JT == JT_UNIVERSAL) { sincos(qu[0], &sc[0], &sc[1]); sincos(qu[1], &sc[2], &sc[3]); }
#pragma unroll
            for (int i = 0; i < lscCount<JT>(); ++i) me.st(LR_SC + i, sc[i]);
        }
        LJoint<JT> k; ljoint<JT, false>(bc, qu, sc, k);
        double* cv = cy + (ph ? LF_V2 : LF_V)*SBK_CARRY_STRIDE;
        const int vbuf = ph ? vw : vr;
        const CacheRefT<BLK> pa = lrecOf<BLK>(c, inst, bc.parentLink);
        const bool fromCarry = (bc.flags & BF_PARENT_PREV) != 0;
        const SV vP = fromCarry ? lcyLoadSV(cv) : pa.ldSV(vbuf);
        SV vJ, cJ; ljointVel<JT>(k, qu + NQ, up, vJ, cJ);
        const SV v = xMotion(k.R, k.p, vP) + vJ;
        lcyStoreSV(cv, v);
        if (ph) {
            if (bc.flags & (BF_STORE_LINK | BF_TIP)) me.stSV(lrV(d) + vw, v);
            break;
        }
        // ---- acceleration sweep of this evaluation (cf. lOutwardBody) ----------------------------------------
        double nu[dim1(d)], upd[dim1(d)], f[dim1(NQ + d)], G[dim1(lgCount<JT>())];
#pragma unroll
        for (int i = 0; i < lgCount<JT>(); ++i) G[i] = pf[(L::G + i)*SBK_CARRY_STRIDE];
#pragma unroll
        for (int j = 0; j < d; ++j) nu[j] = pf[(L::NU + j)*SBK_CARRY_STRIDE];
        const SV aP = fromCarry ? lcyLoadSV(cy + LF_A*SBK_CARRY_STRIDE) : pa.ldSV(6);
        const SV cc = cJ + crossMotion(v, vJ);
        const SV aPlus = xMotion(k.R, k.p, aP);
        const double ap[6] = {aPlus.w.x, aPlus.w.y, aPlus.w.z, aPlus.v.x, aPlus.v.y, aPlus.v.z};
        SV a = aPlus + cc;
        if constexpr (JT == JT_FREE) {
#pragma unroll
            for (int i = 0; i < 6; ++i) upd[i] = nu[i] - ap[i];
            a.w = a.w + mk(upd[0], upd[1], upd[2]); a.v = a.v + mk(upd[3], upd[4], upd[5]);
        } else if constexpr (JT == JT_BALL) {
#pragma unroll
            for (int j = 0; j < 3; ++j) upd[j] = nu[j] - (ap[j] + (G[j]*ap[3] + G[3+j]*ap[4] + G[6+j]*ap[5]));
            a.w = a.w + mk(upd[0], upd[1], upd[2]);
        } else {
#pragma unroll
            for (int i = 0; i < d; ++i) {
                double s = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) s += G[6*i+r]*ap[r];
                upd[i] = nu[i] - s;
            }
            if constexpr (JT == JT_PIN) a.w.z += upd[0];
            else if constexpr (JT == JT_SLIDER) a.v.x += upd[0];
            else { a.w = a.w + mk(k.cb*upd[0], upd[1], k.sb*upd[0]); }      // Universal
        }
        lcyStoreSV(cy + LF_A*SBK_CARRY_STRIDE, a);
        if (bc.flags & BF_STORE_LINK) me.stSV(lrV(d) + 6, a);
        ljointQdot<JT>(qu, qu + NQ, f);
        ljointUdot<JT>(k, up, upd, f + NQ);
        // ---- this body's slots of the next stage state ---------------------------------------------------------
        double err[dim1(NQ + d)], y0u[dim1(d)];
        const bool last = stage == 4;
#pragma unroll
        for (int i = 0; i < NQ + d; ++i) {
            const int slot = i < NQ ? bc.q0 + i : nq + bc.u0 + (i - NQ);
            const double y0 = (stage == 0) ? qu[i] : pf[(L::Y + i)*SBK_CARRY_STRIDE];
            const double f0 = (sg.a0 != 0) ? (L::HAS_F0 != 0 ? pf[(L::F0 + i)*SBK_CARRY_STRIDE] : f0d[i]) : 0.0;
            const double f23 = (sg.m2 != 0 || sg.m3 != 0) ? (L::HAS_F23 != 0 ? pf[(L::F23 + i)*SBK_CARRY_STRIDE] : f23d[i]) : 0.0;
            const double f2 = (sg.m2 != 0) ? f23 : 0.0, f3 = (sg.m3 != 0) ? f23 : 0.0;
            if (sg.fdst) stS<BLK>(c, inst, sg.fdst, slot, f[i]);
            const double r = y0 + sg.hk*(((sg.a0*f0 + sg.m2*f2) + sg.m3*f3) + sg.mf*f[i]);
            if (last) { err[i] = 0.2*fabs(r - qu[i]); stS<BLK>(c, inst, w.Ynext, slot, r); stS<BLK>(c, inst, w.W, slot, err[i]); }
            else stS<BLK>(c, inst, w.W, slot, r);
            qu[i] = r;
            if (i >= NQ) y0u[i - NQ] = y0;
        }
        if (last) {      // error norm sums, IntegratorRep.h:454-513 (cf. rkmErrorNorm: same order, same operands)
            double qAcc = cy[LF_QACC*SBK_CARRY_STRIDE], uAcc = cy[LF_UACC*SBK_CARRY_STRIDE];
#pragma unroll
            for (int i = 0; i < d; ++i) {
                const double a0 = fabs(y0u[i]);
                const double scl = (a0*1.0 > 1.0) ? 1.0/a0 : 1.0;
                uAcc = normAcc(uAcc, scl*err[NQ + i], w.useInfNorm);
            }
            if constexpr (JT == JT_BALL || JT == JT_FREE) {
                double o[4];
                const V3 du = quatNInvTimes(qu, err);
                quatNTimes(qu, du, o);
#pragma unroll
                for (int i = 0; i < 4; ++i) qAcc = normAcc(qAcc, o[i], w.useInfNorm);
#pragma unroll
                for (int i = 4; i < NQ; ++i) qAcc = normAcc(qAcc, err[i], w.useInfNorm);
                const double e = sqrt(qu[0]*qu[0] + qu[1]*qu[1] + qu[2]*qu[2] + qu[3]*qu[3]) - 1.0;
                cy[LF_QUATACC*SBK_CARRY_STRIDE] = normAcc(cy[LF_QUATACC*SBK_CARRY_STRIDE], e, w.useInfNorm);
            } else {
#pragma unroll
                for (int i = 0; i < NQ; ++i) qAcc = normAcc(qAcc, err[i], w.useInfNorm);
            }
            cy[LF_QACC*SBK_CARRY_STRIDE] = qAcc; cy[LF_UACC*SBK_CARRY_STRIDE] = uAcc;
        }
    }
}

// Velocity / inward body steps with an explicit state pointer and velocity buffer (the drivers of sbk_local.cuh use
// Ctx::q/u and buffer 0).
// LOCK (error-controlled kernel): the CTA's threads meet at every body, threads without work (on = false) just keep pace --
// the warps then share their instruction fetches instead of each streaming the kernel from L2 on its own.
template <int JMASK, bool LOCK = false>
SBK_HD void lInwardSweep(const Ctx& c0, const LTables& T, const int inst, double* cy, const double* S, const int vb, const bool on = true) {
    constexpr bool BLK = SBK_DEV_BLK;
    Ctx c = c0; c.q = S; c.u = S + (BLK ? (long long)c.nq*BLK_LANES : (long long)c.nq*c.sStride);
    #define SBK_PFSLOT(par) (cy + (LF_PF + LPF_ROWS*((par) & 1))*SBK_CARRY_STRIDE)
#pragma unroll 1
    for (int b = c.nb; b >= 1; --b) {          // trip b = nb only requests the first body's rows
        if constexpr (LOCK) ctaBarrier();
        if (!on) continue;
        if (b > 1) lPrefetchIn<JMASK, BLK>(c, T.bodies[b - 1], inst, SBK_PFSLOT(b - 1), S, vb);
        lpfCommit(); lpfWait();
        if (b < c.nb) { const LBody& bc = T.bodies[b]; SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lInwardBody<JT>(c, T, bc, inst, cy, vb, SBK_PFSLOT(b)))); }
    }
}
// stage 0..4: acceleration sweep of the evaluation at S + next stage state + its velocity data; stage < 0: velocity data of S only
template <int JMASK, bool LOCK = false>
SBK_HD void lFusedOutSweep(const Ctx& c, const LTables& T, const int inst, double* cy, const LRkmWork& w, const double* S,
                           const int stage, const double h, const int vr, const int vw, const bool on = true) {
    constexpr bool BLK = SBK_DEV_BLK;
    const LStage sg = lstageOf(stage, h, w);
    if (on) {
        SV a0 = zeroSV(); a0.v = mk(-c.gx, -c.gy, -c.gz);
        lcyStoreSV(cy + LF_V*SBK_CARRY_STRIDE, zeroSV()); lcyStoreSV(cy + LF_A*SBK_CARRY_STRIDE, a0); lcyStoreSV(cy + LF_V2*SBK_CARRY_STRIDE, zeroSV());
        if (T.bodies[0].flags & BF_STORE_LINK) {     // Ground's links: v = 0 in both buffers, a = -g (gravity as a base acceleration)
            const CacheRefT<BLK> g = lrecOf<BLK>(c, inst, T.bodies[0].rec);
            g.stSV(lrV(0) + 6, a0); g.stSV(lrV(0) + vw, zeroSV()); g.stSV(lrV(0) + vr, zeroSV());
        }
        if (stage == 4) { cy[LF_QACC*SBK_CARRY_STRIDE] = 0; cy[LF_UACC*SBK_CARRY_STRIDE] = 0; cy[LF_QUATACC*SBK_CARRY_STRIDE] = 0; }
    }
#pragma unroll 1
    for (int b = 0; b < c.nb; ++b) {           // trip b = 0 only requests the first body's rows
        if constexpr (LOCK) ctaBarrier();
        if (!on) continue;
        if (b + 1 < c.nb) lPrefetchOut<JMASK, BLK>(c, T.bodies[b + 1], inst, SBK_PFSLOT(b + 1), w, S, sg);
        lpfCommit(); lpfWait();
        if (b >= 1) { const LBody& bc = T.bodies[b]; SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lFusedOutBody<JT>(c, bc, inst, cy, w, S, sg, vr, vw, SBK_PFSLOT(b)))); }
    }
    #undef SBK_PFSLOT
}
template <int JMASK, bool LOCK = false>
SBK_HD void lVelSweep(const Ctx& c, const LTables& T, const int inst, double* cy, const LRkmWork& w, const double* S, const int vb, const bool on = true) {
    lFusedOutSweep<JMASK, LOCK>(c, T, inst, cy, w, S, -1, 0.0, vb, vb, on);
}

// q part of the error norm from the stored estimate (W) and the new state (after a projection changed both)
template <bool BLK>
SBK_HD double lqErrAcc(const Ctx& c, const LTables& T, const int inst, const LRkmWork& w) {
    double qAcc = 0;
    for (int b = 1; b < c.nb; ++b) {
        const LBody& bc = T.bodies[b];
        int first = 0;
        if (bc.joint == JT_BALL || bc.joint == JT_FREE) {
            double q[4], e[4], o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = ldS<BLK>(c, inst, w.Ynext, bc.q0 + i); e[i] = ldS<BLK>(c, inst, w.W, bc.q0 + i); }
            const V3 du = quatNInvTimes(q, e);
            quatNTimes(q, du, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) qAcc = normAcc(qAcc, o[i], w.useInfNorm);
            first = 4;
        }
        const int nqb = nqOfJoint(bc.joint);
        for (int i = first; i < nqb; ++i) qAcc = normAcc(qAcc, ldS<BLK>(c, inst, w.W, bc.q0 + i), w.useInfNorm);
    }
    return qAcc;
}

// End of an attempt from the error sums of the last outward sweep: error norm, the projection rule of attemptDAEStep
// (AbstractIntegratorRep.cpp:137-208), quaternion normalisation with the radial part of the error estimate removed
// (RigidBodyNodeSpec_Ball.h:417-434) and the recomputed norm (takeOneStep, AbstractIntegratorRep.cpp:556).
template <bool BLK>
SBK_HD RkmStepResult lFinishAttempt(const Ctx& c, const LTables& T, const int inst, const LRkmWork& w, const double qAcc, const double uAcc, const double quatAcc) {
    const int nq = c.nq, nu = c.nu;
    RkmStepResult res; res.projected = 0;
    const double uNorm = w.useInfNorm ? uAcc : (nu ? sqrt(uAcc/nu) : 0.0);
    double qNorm = w.useInfNorm ? qAcc : (nq ? sqrt(qAcc/nq) : 0.0);
    res.errNorm = normMax(uNorm, qNorm);
    if (c.nquat > 0 && !(res.errNorm > 16.0*w.accuracy)) {          // project only if errNorm <= 2^4 * accuracy (:165-166)
        const double quatNorm = w.useInfNorm ? quatAcc : sqrt(quatAcc/c.nquat);
        if (quatNorm > projectionLimit(w.consTol)) res.errNorm = 1.0/0.0;      // convergence failure: too far off the manifold to project
        else if (quatNorm > w.consTol || w.projectEveryStep) {
            for (int b = 1; b < c.nb; ++b) {
                const LBody& bc = T.bodies[b];
                if (bc.joint != JT_BALL && bc.joint != JT_FREE) continue;
                double q[4], e[4], n2 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = ldS<BLK>(c, inst, w.Ynext, bc.q0 + i); e[i] = ldS<BLK>(c, inst, w.W, bc.q0 + i); n2 += q[i]*q[i]; }
                const double n = sqrt(n2);
                double dt = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = q[i]/n; dt += e[i]*q[i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { stS<BLK>(c, inst, w.Ynext, bc.q0 + i, q[i]); stS<BLK>(c, inst, w.W, bc.q0 + i, e[i] - dt*q[i]); }
            }
            res.projected = 1;
            const double qAcc2 = lqErrAcc<BLK>(c, T, inst, w);        // the u part does not change
            qNorm = w.useInfNorm ? qAcc2 : (nq ? sqrt(qAcc2/nq) : 0.0);
            res.errNorm = normMax(uNorm, qNorm);
        }
    }
    return res;
}

// Per-instance state of the fused integrator that survives from one attempt to the next.
struct LRkmState { int vb; bool velValid; };    // velocity buffer (row offset 0 / LR_VBUF) with the data of the state in Y; whether it is current

// One RKM attempt of size h from the state in Y.  fresh = false: retry of a failed attempt (Y, F0 kept; a smaller h).
// On return y1 is in Ynext, the error estimate in W, and the velocity data of y1 are in buffer st.vb (st.velValid)
// -- valid for the next step if the caller accepts y1 (and then makes Ynext the new Y).
template <int JMASK>
SBK_HD RkmStepResult lRkmAttempt(const Ctx& c, const LTables& T, const int inst, const LRkmWork& w, const double h, double* cy,
                                 LRkmState& st, const bool fresh = true) {
    constexpr bool BLK = SBK_DEV_BLK;
    const int nq = c.nq, nu = c.nu, ny = nq + nu;
    int vb = st.vb;        // row offset of the velocity buffer in use: 0 or LR_VBUF
    if (fresh) {
        if (!st.velValid) lVelSweep<JMASK>(c, T, inst, cy, w, w.Y, vb);
        lInwardSweep<JMASK>(c, T, inst, cy, w.Y, vb);
        lFusedOutSweep<JMASK>(c, T, inst, cy, w, w.Y, 0, h, vb, vb ^ LR_VBUF); vb ^= LR_VBUF;
    } else {       // W = y0 + h/3 f0 with the new h, and its velocity data
        for (int i = 0; i < ny; ++i) stS<BLK>(c, inst, w.W, i, ldS<BLK>(c, inst, w.Y, i) + (h/3)*ldS<BLK>(c, inst, w.F0, i));
        lVelSweep<JMASK>(c, T, inst, cy, w, w.W, vb);
    }
#pragma unroll 1
    for (int stage = 1; stage < 5; ++stage) {
        lInwardSweep<JMASK>(c, T, inst, cy, w.W, vb);
        lFusedOutSweep<JMASK>(c, T, inst, cy, w, w.W, stage, h, vb, vb ^ LR_VBUF); vb ^= LR_VBUF;
    }
    st.vb = vb; st.velValid = true;
    const RkmStepResult res = lFinishAttempt<BLK>(c, T, inst, w, cy[LF_QACC*SBK_CARRY_STRIDE], cy[LF_UACC*SBK_CARRY_STRIDE], cy[LF_QUATACC*SBK_CARRY_STRIDE]);
    if (res.projected) st.velValid = false;                            // the quaternions moved: velocity data must be redone
    return res;
}

// The same attempt for a CTA that votes as a whole (error-controlled kernel): every thread walks every sweep some thread of the
// CTA needs (the first-evaluation sweeps of threads starting a step, the restart sweep of threads retrying one), threads
// without work in a sweep (or without work at all: mine = false) keep pace at the per-body barriers.
template <int JMASK, class VOTE>
SBK_HD RkmStepResult lRkmAttemptLockstep(const Ctx& c, const LTables& T, const int inst, const LRkmWork& w, const double h, double* cy,
                                         LRkmState& st, const bool fresh, const bool mine, const VOTE& vote) {
    constexpr bool BLK = SBK_DEV_BLK;
    const int ny = c.nq + c.nu;
    int vb = st.vb;
    const bool doFresh = mine && fresh, doRetry = mine && !fresh, needVel = doFresh && !st.velValid;
    if (vote(needVel)) lVelSweep<JMASK, true>(c, T, inst, cy, w, w.Y, vb, needVel);
    if (vote(doFresh)) {
        lInwardSweep<JMASK, true>(c, T, inst, cy, w.Y, vb, doFresh);
        lFusedOutSweep<JMASK, true>(c, T, inst, cy, w, w.Y, 0, h, vb, vb ^ LR_VBUF, doFresh);
    }
    if (doFresh) vb ^= LR_VBUF;
    if (vote(doRetry)) {
        if (doRetry) for (int i = 0; i < ny; ++i) stS<BLK>(c, inst, w.W, i, ldS<BLK>(c, inst, w.Y, i) + (h/3)*ldS<BLK>(c, inst, w.F0, i));
        lVelSweep<JMASK, true>(c, T, inst, cy, w, w.W, vb, doRetry);
    }
#pragma unroll 1
    for (int stage = 1; stage < 5; ++stage) {
        lInwardSweep<JMASK, true>(c, T, inst, cy, w.W, vb, mine);
        lFusedOutSweep<JMASK, true>(c, T, inst, cy, w, w.W, stage, h, vb, vb ^ LR_VBUF, mine); vb ^= LR_VBUF;
    }
    RkmStepResult res; res.errNorm = 0; res.projected = 0;
    if (mine) {
        st.vb = vb; st.velValid = true;
        res = lFinishAttempt<BLK>(c, T, inst, w, cy[LF_QACC*SBK_CARRY_STRIDE], cy[LF_UACC*SBK_CARRY_STRIDE], cy[LF_QUATACC*SBK_CARRY_STRIDE]);
        if (res.projected) st.velValid = false;
    }
    return res;
}

// Error-controlled stepping with the fused attempt (cf. tpiRkmAdaptive): y1 goes to a second state buffer and the two
// swap roles when a step is accepted.  On return w.Y points at the buffer holding the advanced state.
template <int JMASK, class VOTE = WarpVote>
SBK_HD void lRkmAdaptive(const Ctx& c, const LTables& T, const int inst, LRkmWork& w, const StepLimits& lim, const double tFinal,
                         const int allowInterpolation, const int maxAttempts, AdaptiveState& ast, double* cy, LRkmState& st,
                         double& lastErr, int& nproj, const bool live = true, const VOTE vote = VOTE()) {
    int budget = maxAttempts; bool fresh = true;
    for (;;) {                                   // one attempt per trip; see WarpVote (sbk_rkm.cuh)
        const bool mine = live && ast.t < tFinal && budget > 0;
        if (!vote(mine)) break;
        bool limited = false; double t1 = ast.t;
        if (mine) {
            if (allowInterpolation) t1 = ast.t + ast.h;
            else if (tFinal < ast.t + 0.95*ast.h)  { limited = true; t1 = tFinal; }
            else if (tFinal > ast.t + 1.001*ast.h) t1 = ast.t + ast.h;
            else t1 = tFinal;
        }
        RkmStepResult r; r.errNorm = 0; r.projected = 0;
        if constexpr (VOTE::CTA) r = lRkmAttemptLockstep<JMASK>(c, T, inst, w, t1 - ast.t, cy, st, fresh, mine, vote);
        else if (mine) r = lRkmAttempt<JMASK>(c, T, inst, w, t1 - ast.t, cy, st, fresh);
        if (mine) {
            ++ast.attempts; --budget; nproj += r.projected; lastErr = r.errNorm;
            fresh = adjustStepSize(r.errNorm, lim, limited, ast.h);
            if (fresh) {                                  // accept: y1 becomes the state
                double* t = w.Y; w.Y = w.Ynext; w.Ynext = t;
                ast.lastStep = t1 - ast.t; ast.t = t1; ++ast.steps;
            }
        }
    }
    if (!fresh) st.velValid = false;             // out of budget inside a failing step: Y still holds y0
}

} // namespace sbkd
