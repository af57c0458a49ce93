// sbk_sweeps.cuh -- per-body steps of the O(n) sweeps and the per-instance drivers.
//
// One *work item* runs these functions for one (instance, body) pair.  The functions are
// written once and used by every execution plan (thread-per-instance, register-resident
// fused, level-parallel); plans differ only in the Cache/State accessor types they pass in
// and in who loops over bodies.
//
// Sweeps (reference file:line each one replaces)
//   kinBody      sweeps A+B, base->tip   RigidBodyNodeSpec.h:229-333, RigidBodyNodeSpec.cpp:44-129,
//                                        RigidBodyNode.cpp:54-174, RigidBodyNodeSpec_{Pin,Slider,
//                                        Universal,Ball,Free}.h
//   inwardBody   sweeps C(+bias)+D, tip->base
//                                        RigidBodyNodeSpec.cpp:249-325 (ABI), RigidBodyNode.cpp:201-213
//                                        (P*a+b), :355-400 (z, eps, zPlus), :483-515 (M^-1 pass 1),
//                                        Force_Gravity.cpp:514-569, Force.cpp:339-351,434-443
//   outwardBody  sweep E, base->tip      RigidBodyNodeSpec.cpp:408-446, :521-552, calcQDotDot
//   mulMOut/In, residOut/In              RigidBodyNodeSpec.cpp:566-695
//
// Per-body cache record (doubles), dof = d:
//   XGB 12 | VGB 6 | L 3 | MK 9 (c_G, G_G) | ACOR 6 | GYRO 6 | ZB 6 | PPLUS 21 | ZPLUS 6 | AGB 6 |
//   H 6d | G 6d | DI d*d | EPS d
#pragma once
#include "sbk_math.cuh"

namespace sbkd {

enum { JT_GROUND = 0, JT_PIN = 1, JT_SLIDER = 2, JT_UNIVERSAL = 3, JT_BALL = 4, JT_FREE = 5 };
enum { FK_SPRING = 2, FK_DAMPER = 3 };

template <int JT> struct JointDims;
template <> struct JointDims<JT_PIN>       { enum { nq = 1, nu = 1 }; };
template <> struct JointDims<JT_SLIDER>    { enum { nq = 1, nu = 1 }; };
template <> struct JointDims<JT_UNIVERSAL> { enum { nq = 2, nu = 2 }; };
template <> struct JointDims<JT_BALL>      { enum { nq = 4, nu = 3 }; };
template <> struct JointDims<JT_FREE>      { enum { nq = 7, nu = 6 }; };

// cache record layout
enum { F_XGB = 0, F_VGB = 12, F_L = 18, F_MK = 21, F_ACOR = 30, F_GYRO = 36, F_ZB = 42,
       F_PPLUS = 48, F_ZPLUS = 69, F_AGB = 75, F_H = 81 };
SBK_HD constexpr int fG(int d)   { return F_H + 6*d; }
SBK_HD constexpr int fDI(int d)  { return F_H + 12*d; }
SBK_HD constexpr int fEPS(int d) { return F_H + 12*d + d*d; }
SBK_HD constexpr int cacheRecordSize(int d) { return F_H + 13*d + d*d; }
enum { CACHE_RECORD_MAX = F_H + 13*6 + 36 };   // 195

// Per-body constants (batch-shared; staged into shared memory by the kernels).
struct BodyConst {
    double X_PF[12];        // R_PF row-major, p_PF
    double X_MB[12];        // ~X_BM (RigidBodyNode.h:1012)
    double mass;
    double com_B[3];
    double G_B[6];          // unit inertia about Bo in B: xx yy zz xy xz yz
    long long cacheBase;    // element offset of this body's cache record (plan dependent)
    long long parentCacheBase;
    int joint, parent, q0, u0;
    int nchild, childStart, quat, nforce;
    int forceStart, level, pad0, pad1;
};
struct ForceConst { int kind; int coord; double a; double b; };

// What one work item sees.  q/u/qdot/... are SoA arrays addressed [slot*sStride + sOff].
struct Ctx {
    const BodyConst*  bodies;
    const int*        children;
    const ForceConst* forces;
    int nb, nq, nu, nquat;
    double gx, gy, gz;              // gravity vector g*d (Force_Gravity.cpp:532), 0 if none
    // cache addressing: element (body record base B, field k) at cache[B + k*cStride + cOff]
    double*   cache;
    long long cStride, cOff;
    // state addressing
    long long sStride, sOff;
    const double* q; const double* u;
    double* qdot; double* udot; double* qdotdot; double* qerr;
    // operator inputs / outputs (nullable)
    const double* fmobIn;  const double* FbodyIn;   // applied forces for operator forms
    double* fmobOut; double* FbodyOut;              // force-subsystem results (getter)
    const double* vecIn; double* vecOut;            // generic nu-vectors for M, M^-1, residual
    int* status;                                    // per-instance status word
};

struct CacheRef {     // accessor for one body's record
    double* p; long long stride;
    SBK_HD double ld(int k) const { return p[(long long)k*stride]; }
    SBK_HD void   st(int k, double v) const { p[(long long)k*stride] = v; }
    SBK_HD V3 ld3(int k) const { return mk(ld(k), ld(k+1), ld(k+2)); }
    SBK_HD void st3(int k, V3 v) const { st(k, v.x); st(k+1, v.y); st(k+2, v.z); }
    SBK_HD SV ldSV(int k) const { SV r; r.w = ld3(k); r.v = ld3(k+3); return r; }
    SBK_HD void stSV(int k, SV v) const { st3(k, v.w); st3(k+3, v.v); }
    SBK_HD M3 ldM3(int k) const { M3 R;
#pragma unroll
        for (int i = 0; i < 9; ++i) R.a[i] = ld(k+i); return R; }
    SBK_HD void stM3(int k, const M3& R) const {
#pragma unroll
        for (int i = 0; i < 9; ++i) st(k+i, R.a[i]); }
    SBK_HD S3 ldS3(int k) const { S3 s; s.xx = ld(k); s.yy = ld(k+1); s.zz = ld(k+2); s.xy = ld(k+3); s.xz = ld(k+4); s.yz = ld(k+5); return s; }
    SBK_HD void stS3(int k, const S3& s) const { st(k, s.xx); st(k+1, s.yy); st(k+2, s.zz); st(k+3, s.xy); st(k+4, s.xz); st(k+5, s.yz); }
    SBK_HD ABI ldABI(int k) const { ABI P; P.M = ldS3(k); P.J = ldS3(k+6); P.F = ldM3(k+12); return P; }
    SBK_HD void stABI(int k, const ABI& P) const { stS3(k, P.M); stS3(k+6, P.J); stM3(k+12, P.F); }
};
SBK_HD CacheRef cacheOf(const Ctx& c, long long base) { CacheRef r; r.p = c.cache + base + c.cOff; r.stride = c.cStride; return r; }
SBK_HD double ldS(const Ctx& c, const double* a, int slot) { return a[(long long)slot*c.sStride + c.sOff]; }
SBK_HD void   stS(const Ctx& c, double* a, int slot, double v) { a[(long long)slot*c.sStride + c.sOff] = v; }

SBK_HD M3 loadR(const double* X) { M3 R;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.a[i] = X[i]; return R; }
SBK_HD V3 loadP(const double* X) { return mk(X[9], X[10], X[11]); }

// Unnormalised N(q) * w for a scalar-first quaternion (Rotation.h:712-720).
SBK_HD void quatNTimes(const double* q, V3 w, double* out) {
    const double e0 = q[0]/2, e1 = q[1]/2, e2 = q[2]/2, e3 = q[3]/2;
    const double ne1 = -e1, ne2 = -e2, ne3 = -e3;
    out[0] = ne1*w.x + ne2*w.y + ne3*w.z;
    out[1] = e0*w.x  + e3*w.y  + ne2*w.z;
    out[2] = ne3*w.x + e0*w.y  + e1*w.z;
    out[3] = e2*w.x  + ne1*w.y + e0*w.z;
}
// Unnormalised NInv(q) * qd (Rotation.h:742-748)
SBK_HD V3 quatNInvTimes(const double* q, const double* qd) {
    const double e0 = 2*q[0], e1 = 2*q[1], e2 = 2*q[2], e3 = 2*q[3];
    const double ne1 = -e1, ne2 = -e2, ne3 = -e3;
    return mk(ne1*qd[0] + e0*qd[1] + ne3*qd[2] + e2*qd[3],
              ne2*qd[0] + e3*qd[1] + e0*qd[2]  + ne1*qd[3],
              ne3*qd[0] + ne2*qd[1] + e1*qd[2] + e0*qd[3]);
}

//==============================================================================================
// Sweeps A+B for one body (base->tip).  Needs the parent's XGB/VGB already in the cache.
//==============================================================================================
template <int JT>
SBK_HDN void kinBody(const Ctx& c, const BodyConst& bc) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    const CacheRef me = cacheOf(c, bc.cacheBase), pa = cacheOf(c, bc.parentCacheBase);

    double q[NQ], u[d];
#pragma unroll
    for (int i = 0; i < NQ; ++i) q[i] = ldS(c, c.q, bc.q0 + i);
#pragma unroll
    for (int i = 0; i < d; ++i)  u[i] = ldS(c, c.u, bc.u0 + i);

    // ---- mobilizer-specific: X_FM, H_FM, HDot_FM (all expressed in F) ----------------------
    M3 R_FM; V3 p_FM = zero3();
    V3 Hw[d], Hv[d];          // H_FM columns (angular, linear)
    V3 HDw[d];                // HDot_FM angular part (linear part is zero for all five mobilizers)
#pragma unroll
    for (int j = 0; j < d; ++j) { Hw[j] = zero3(); Hv[j] = zero3(); HDw[j] = zero3(); }

    if constexpr (JT == JT_PIN) {                 // RigidBodyNodeSpec_Pin.h:103-140
        double s, co; sincos(q[0], &s, &co);
        R_FM.a[0] = co; R_FM.a[1] = -s; R_FM.a[2] = 0;
        R_FM.a[3] = s;  R_FM.a[4] = co; R_FM.a[5] = 0;
        R_FM.a[6] = 0;  R_FM.a[7] = 0;  R_FM.a[8] = 1;
        Hw[0] = mk(0, 0, 1);
    } else if constexpr (JT == JT_SLIDER) {       // RigidBodyNodeSpec_Slider.h:91-127
        R_FM = identity3(); p_FM = mk(q[0], 0, 0);
        Hv[0] = mk(1, 0, 0);
    } else if constexpr (JT == JT_UNIVERSAL) {    // RigidBodyNodeSpec_Universal.h:124-192, Rotation.cpp:241-264
        double s1, c1, s2, c2; sincos(q[0], &s1, &c1); sincos(q[1], &s2, &c2);
        R_FM.a[0] = c2;      R_FM.a[1] = 0;  R_FM.a[2] = s2;
        R_FM.a[3] = s2*s1;   R_FM.a[4] = c1; R_FM.a[5] = -s1*c2;
        R_FM.a[6] = -s2*c1;  R_FM.a[7] = s1; R_FM.a[8] = c1*c2;
        Hw[0] = mk(1, 0, 0);
        Hw[1] = col(R_FM, 1);
    } else {                                      // Ball / Free: RigidBodyNodeSpec_Ball.h:113-180, _Free.h:142-222
        const double quatLen = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
        if (c.qerr) stS(c, c.qerr, bc.quat, quatLen - 1.0);
        const double oon = 1.0/quatLen;
        R_FM = rotFromQuat(q[0]*oon, q[1]*oon, q[2]*oon, q[3]*oon);
        Hw[0] = mk(1, 0, 0); Hw[1] = mk(0, 1, 0); Hw[2] = mk(0, 0, 1);
        if constexpr (JT == JT_FREE) {
            p_FM = mk(q[4], q[5], q[6]);
            Hv[3] = mk(1, 0, 0); Hv[4] = mk(0, 1, 0); Hv[5] = mk(0, 0, 1);
        }
    }

    // ---- calcBodyTransforms (RigidBodyNodeSpec.h:554-569) ------------------------------------
    const M3 R_PF = loadR(bc.X_PF), R_MB = loadR(bc.X_MB);
    const V3 p_PF = loadP(bc.X_PF), p_MB = loadP(bc.X_MB);
    const M3 R_GP = pa.ldM3(F_XGB); const V3 p_GP = pa.ld3(F_XGB + 9);
    const SV V_GP = pa.ldSV(F_VGB);

    const V3 r     = mul(R_FM, p_MB);                 // r_MB_F = R_FM * p_MB
    const M3 R_FB  = mul(R_FM, R_MB);  const V3 p_FB = p_FM + r;
    const M3 R_PB  = mul(R_PF, R_FB);  const V3 p_PB = p_PF + mul(R_PF, p_FB);
    const M3 R_GB  = mul(R_GP, R_PB);
    const V3 l     = mul(R_GP, p_PB);                 // Phi: p_PB_G (RigidBodyNode.cpp:61)
    const V3 p_GB  = p_GP + l;
    me.stM3(F_XGB, R_GB); me.st3(F_XGB + 9, p_GB); me.st3(F_L, l);

    // ---- H = R_GF (H_FM + H_MB_F)  (RigidBodyNodeSpec.cpp:44-74) -----------------------------
    const M3 R_GF = mul(R_GP, R_PF);
    SV H[d];
#pragma unroll
    for (int j = 0; j < d; ++j) {
        H[j].w = mul(R_GF, Hw[j]);
        H[j].v = mul(R_GF, Hv[j] + cross(Hw[j], r));  // H_MB_F[1] = -r x H_FM[0]
        me.stSV(F_H + 6*j, H[j]);
    }

    // ---- mass properties in Ground (RigidBodyNode.cpp:54-84) ---------------------------------
    S3 G_B; G_B.xx = bc.G_B[0]; G_B.yy = bc.G_B[1]; G_B.zz = bc.G_B[2]; G_B.xy = bc.G_B[3]; G_B.xz = bc.G_B[4]; G_B.yz = bc.G_B[5];
    const S3 G_G = reexpressSym(R_GB, G_B);
    const V3 c_G = mul(R_GB, mk(bc.com_B[0], bc.com_B[1], bc.com_B[2]));
    me.st3(F_MK, c_G); me.stS3(F_MK + 3, G_G);

    // ---- velocity (RigidBodyNodeSpec.h:305-333) ------------------------------------------------
    V3 w_FM = zero3(); SV V_PB = zeroSV();
#pragma unroll
    for (int j = 0; j < d; ++j) { w_FM = w_FM + u[j]*Hw[j]; V_PB = V_PB + u[j]*H[j]; }
    if constexpr (JT == JT_UNIVERSAL) HDw[1] = cross(w_FM, col(R_FM, 1));   // _Universal.h:176-190

    // HDot (RigidBodyNodeSpec.cpp:82-129) and VD = HDot*u
    const V3 w_GP = V_GP.w, v_GP = V_GP.v;
    const V3 wxr = cross(w_FM, r);
    SV VD = zeroSV();
#pragma unroll
    for (int j = 0; j < d; ++j) {
        SV HD;
        HD.w = mul(R_GF, HDw[j]) + cross(w_GP, H[j].w);
        HD.v = mul(R_GF, cross(HDw[j], r) + cross(Hw[j], wxr)) + cross(w_GP, H[j].v);
        VD = VD + u[j]*HD;
    }

    // ---- joint-independent velocity kinematics (RigidBodyNode.cpp:97-174) ----------------------
    const SV V_GB = phiT(l, V_GP) + V_PB;
    me.stSV(F_VGB, V_GB);
    const V3 w = V_GB.w;
    SV gyro; gyro.w = bc.mass*cross(w, mul(G_G, w)); gyro.v = bc.mass*cross(w, cross(w, c_G));
    me.stSV(F_GYRO, gyro);
    SV acor; acor.w = VD.w; acor.v = VD.v + cross(w_GP, V_GB.v - v_GP);
    me.stSV(F_ACOR, acor);

    // ---- qdot = N(q) u  ------------------------------------------------------------------------
    if (c.qdot) {
        if constexpr (JT == JT_BALL || JT == JT_FREE) {
            double qd[4]; quatNTimes(q, mk(u[0], u[1], u[2]), qd);
#pragma unroll
            for (int i = 0; i < 4; ++i) stS(c, c.qdot, bc.q0 + i, qd[i]);
            if constexpr (JT == JT_FREE) {
#pragma unroll
                for (int i = 0; i < 3; ++i) stS(c, c.qdot, bc.q0 + 4 + i, u[3+i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < d; ++i) stS(c, c.qdot, bc.q0 + i, u[i]);
        }
    }
}

//==============================================================================================
// Inward body step.  MODE bits select what is done:
//   IN_ABI    articulated-body inertia: P, D, DI, G, PPlus (+ ZB = P*a + b)
//   IN_Z      residual pass: z, eps, zPlus
//   IN_BIAS   z starts from ZB (forward dynamics); otherwise from 0 (M^-1)
//   IN_FORCES applied forces come from the lowered force elements (gravity/spring/damper)
//             rather than from c.fmobIn / c.FbodyIn
//==============================================================================================
enum { IN_ABI = 1, IN_Z = 2, IN_BIAS = 4, IN_FORCES = 8 };

template <int JT, int MODE>
SBK_HDN void inwardBody(const Ctx& c, const BodyConst& bc, const int bodyIndex) {
    constexpr int d = JointDims<JT>::nu;
    const CacheRef me = cacheOf(c, bc.cacheBase);

    SV H[d];
#pragma unroll
    for (int j = 0; j < d; ++j) H[j] = me.ldSV(F_H + 6*j);
    SV G[d]; SV zb = zeroSV();
    const V3 c_G = me.ld3(F_MK);

    if constexpr ((MODE & IN_ABI) != 0) {
        // ---- realizeArticulatedBodyInertiasInward (RigidBodyNodeSpec.cpp:249-325) -------------
        const S3 G_G = me.ldS3(F_MK + 3);
        ABI P = abiFromRigid(bc.mass, c_G, G_G);
        for (int k = 0; k < bc.nchild; ++k) {
            const BodyConst& cb = c.bodies[c.children[bc.childStart + k]];
            const CacheRef ch = cacheOf(c, cb.cacheBase);
            addInto(P, shiftABI(ch.ldABI(F_PPLUS), ch.ld3(F_L)));
        }
        SV PH[d];
#pragma unroll
        for (int j = 0; j < d; ++j) PH[j] = mul(P, H[j]);
        double D[d*d], DI[d*d];
#pragma unroll
        for (int i = 0; i < d; ++i)
#pragma unroll
            for (int j = 0; j < d; ++j) D[d*i+j] = dot(H[i].w, PH[j].w) + dot(H[i].v, PH[j].v);
        const bool ok = Inv<d>::run(D, DI);
        if (!ok && c.status) *c.status |= 2;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            SV g = zeroSV();
#pragma unroll
            for (int k = 0; k < d; ++k) g = g + DI[d*k+j]*PH[k];
            G[j] = g; me.stSV(fG(d) + 6*j, g);
        }
#pragma unroll
        for (int i = 0; i < d*d; ++i) me.st(fDI(d) + i, DI[i]);
        // PPlus = P - G*~PH, symmetrised (RigidBodyNodeSpec.cpp:309-324)
        double mm[9], ms[9], in[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { mm[i] = 0; ms[i] = 0; in[i] = 0; }
#pragma unroll
        for (int k = 0; k < d; ++k) {
            const double gw[3] = {G[k].w.x, G[k].w.y, G[k].w.z}, gv[3] = {G[k].v.x, G[k].v.y, G[k].v.z};
            const double pw[3] = {PH[k].w.x, PH[k].w.y, PH[k].w.z}, pv[3] = {PH[k].v.x, PH[k].v.y, PH[k].v.z};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    mm[3*i+j] += gw[i]*pv[j];    // massMoment = G.row(0)*~PH.row(1)
                    ms[3*i+j] += gv[i]*pv[j];    // mass       = G.row(1)*~PH.row(1)
                    in[3*i+j] += gw[i]*pw[j];    // inertia    = G.row(0)*~PH.row(0)
                }
        }
        ABI PP;
        PP.M.xx = P.M.xx - ms[0]; PP.M.yy = P.M.yy - ms[4]; PP.M.zz = P.M.zz - ms[8];
        PP.M.xy = P.M.xy - (ms[3]+ms[1])/2; PP.M.xz = P.M.xz - (ms[6]+ms[2])/2; PP.M.yz = P.M.yz - (ms[7]+ms[5])/2;
        PP.J.xx = P.J.xx - in[0]; PP.J.yy = P.J.yy - in[4]; PP.J.zz = P.J.zz - in[8];
        PP.J.xy = P.J.xy - (in[3]+in[1])/2; PP.J.xz = P.J.xz - (in[6]+in[2])/2; PP.J.yz = P.J.yz - (in[7]+in[5])/2;
#pragma unroll
        for (int i = 0; i < 9; ++i) PP.F.a[i] = P.F.a[i] - mm[i];
        me.stABI(F_PPLUS, PP);
        // realizeArticulatedBodyVelocityCache (RigidBodyNode.cpp:201-213): P*a + b
        zb = mul(P, me.ldSV(F_ACOR)) + me.ldSV(F_GYRO);
        me.stSV(F_ZB, zb);
    } else {
#pragma unroll
        for (int j = 0; j < d; ++j) G[j] = me.ldSV(fG(d) + 6*j);
        if constexpr ((MODE & IN_BIAS) != 0) zb = me.ldSV(F_ZB);
    }

    if constexpr ((MODE & IN_Z) != 0) {
        // ---- applied forces -------------------------------------------------------------------
        SV F = zeroSV(); double f[d];
#pragma unroll
        for (int j = 0; j < d; ++j) f[j] = 0;
        if constexpr ((MODE & IN_FORCES) != 0) {
            // Force::Gravity (Force_Gravity.cpp:538-556): F = (p_CB_G x m g, m g)
            const V3 Fc = bc.mass*mk(c.gx, c.gy, c.gz);
            F.w = cross(c_G, Fc); F.v = Fc;
            // MobilityLinearSpring / Damper in force-index order (Force.cpp:339-351,434-443)
            for (int k = 0; k < bc.nforce; ++k) {
                const ForceConst fc = c.forces[bc.forceStart + k];
                double frc;
                if (fc.kind == FK_SPRING) frc = -fc.a*(ldS(c, c.q, bc.q0 + fc.coord) - fc.b);
                else                      frc = -fc.a*ldS(c, c.u, bc.u0 + fc.coord);
#pragma unroll
                for (int j = 0; j < d; ++j) if (j == fc.coord) f[j] += frc;
            }
            if (c.fmobOut) {
#pragma unroll
                for (int j = 0; j < d; ++j) stS(c, c.fmobOut, bc.u0 + j, f[j]);
            }
            if (c.FbodyOut) {
                stS(c, c.FbodyOut, 6*bodyIndex+0, F.w.x); stS(c, c.FbodyOut, 6*bodyIndex+1, F.w.y); stS(c, c.FbodyOut, 6*bodyIndex+2, F.w.z);
                stS(c, c.FbodyOut, 6*bodyIndex+3, F.v.x); stS(c, c.FbodyOut, 6*bodyIndex+4, F.v.y); stS(c, c.FbodyOut, 6*bodyIndex+5, F.v.z);
            }
        } else {
            if (c.fmobIn) {
#pragma unroll
                for (int j = 0; j < d; ++j) f[j] = ldS(c, c.fmobIn, bc.u0 + j);
            }
            if (c.FbodyIn) {
                F.w = mk(ldS(c, c.FbodyIn, 6*bodyIndex+0), ldS(c, c.FbodyIn, 6*bodyIndex+1), ldS(c, c.FbodyIn, 6*bodyIndex+2));
                F.v = mk(ldS(c, c.FbodyIn, 6*bodyIndex+3), ldS(c, c.FbodyIn, 6*bodyIndex+4), ldS(c, c.FbodyIn, 6*bodyIndex+5));
            }
        }
        // ---- calcUDotPass1Inward (RigidBodyNodeSpec.cpp:355-400) / M^-1 pass 1 (:483-515) ------
        SV z;
        if constexpr ((MODE & IN_BIAS) != 0) z = zb - F; else z = zeroSV();
        for (int k = 0; k < bc.nchild; ++k) {
            const BodyConst& cb = c.bodies[c.children[bc.childStart + k]];
            const CacheRef ch = cacheOf(c, cb.cacheBase);
            z = z + phi(ch.ld3(F_L), ch.ldSV(F_ZPLUS));
        }
        SV zPlus = z;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            const double eps = f[j] - (dot(H[j].w, z.w) + dot(H[j].v, z.v));
            me.st(fEPS(d) + j, eps);
            f[j] = eps;
        }
        SV Ge = zeroSV();
#pragma unroll
        for (int j = 0; j < d; ++j) Ge = Ge + f[j]*G[j];
        zPlus = zPlus + Ge;
        me.stSV(F_ZPLUS, zPlus);
    }
}

//==============================================================================================
// Sweep E for one body (base->tip): udot, A_GB, qdotdot.
//   WITH_COR: add the mobilizer coriolis acceleration (forward dynamics) or not (M^-1).
//==============================================================================================
template <int JT, bool WITH_COR>
SBK_HDN void outwardBody(const Ctx& c, const BodyConst& bc, double* udotDst, double* qdotdotDst) {
    constexpr int d = JointDims<JT>::nu;
    const CacheRef me = cacheOf(c, bc.cacheBase), pa = cacheOf(c, bc.parentCacheBase);
    const SV APlus = phiT(me.ld3(F_L), pa.ldSV(F_AGB));
    double eps[d], udot[d];
#pragma unroll
    for (int j = 0; j < d; ++j) eps[j] = me.ld(fEPS(d) + j);
    SV A = APlus;
    SV Hu = zeroSV();
#pragma unroll
    for (int i = 0; i < d; ++i) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < d; ++j) s += me.ld(fDI(d) + d*i + j)*eps[j];
        const SV Gi = me.ldSV(fG(d) + 6*i);
        udot[i] = s - (dot(Gi.w, APlus.w) + dot(Gi.v, APlus.v));
        Hu = Hu + udot[i]*me.ldSV(F_H + 6*i);
        if (udotDst) stS(c, udotDst, bc.u0 + i, udot[i]);
    }
    A = A + Hu;
    if constexpr (WITH_COR) A = A + me.ldSV(F_ACOR);
    me.stSV(F_AGB, A);

    if (qdotdotDst) {   // calcQDotDot (RigidBodyNodeSpec_Ball.h:355-390, _Free.h:418-455)
        if constexpr (JT == JT_BALL || JT == JT_FREE) {
            double q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = ldS(c, c.q, bc.q0 + i);
            const V3 w = mk(ldS(c, c.u, bc.u0), ldS(c, c.u, bc.u0 + 1), ldS(c, c.u, bc.u0 + 2));
            double Nb[4]; quatNTimes(q, mk(udot[0], udot[1], udot[2]), Nb);
            const double k = -0.25*dot(w, w);
#pragma unroll
            for (int i = 0; i < 4; ++i) stS(c, qdotdotDst, bc.q0 + i, Nb[i] + k*q[i]);
            if constexpr (JT == JT_FREE) {
#pragma unroll
                for (int i = 0; i < 3; ++i) stS(c, qdotdotDst, bc.q0 + 4 + i, udot[3+i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < d; ++i) stS(c, qdotdotDst, bc.q0 + i, udot[i]);
        }
    }
}

//==============================================================================================
// multiplyByM / inverse dynamics passes (RigidBodyNodeSpec.cpp:566-695).
//   WITH_VEL: residual form (adds coriolis a, gyroscopic b, applied forces).
// The outward pass stores A in AGB; the inward pass stores F in ZPLUS.
//==============================================================================================
template <int JT, bool WITH_VEL>
SBK_HDN void idOutBody(const Ctx& c, const BodyConst& bc) {
    constexpr int d = JointDims<JT>::nu;
    const CacheRef me = cacheOf(c, bc.cacheBase), pa = cacheOf(c, bc.parentCacheBase);
    SV A = phiT(me.ld3(F_L), pa.ldSV(F_AGB));
    SV Hu = zeroSV();
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double v = c.vecIn ? ldS(c, c.vecIn, bc.u0 + j) : 0.0;
        Hu = Hu + v*me.ldSV(F_H + 6*j);
    }
    A = A + Hu;
    if constexpr (WITH_VEL) A = A + me.ldSV(F_ACOR);
    me.stSV(F_AGB, A);
}
template <int JT, bool WITH_VEL>
SBK_HDN void idInBody(const Ctx& c, const BodyConst& bc, const int bodyIndex) {
    constexpr int d = JointDims<JT>::nu;
    const CacheRef me = cacheOf(c, bc.cacheBase);
    const V3 c_G = me.ld3(F_MK); const S3 G_G = me.ldS3(F_MK + 3);
    SV F = mulSpatialInertia(bc.mass, c_G, G_G, me.ldSV(F_AGB));
    if constexpr (WITH_VEL) {
        F = F + me.ldSV(F_GYRO);
        if (c.FbodyIn) {
            SV Fa;
            Fa.w = mk(ldS(c, c.FbodyIn, 6*bodyIndex+0), ldS(c, c.FbodyIn, 6*bodyIndex+1), ldS(c, c.FbodyIn, 6*bodyIndex+2));
            Fa.v = mk(ldS(c, c.FbodyIn, 6*bodyIndex+3), ldS(c, c.FbodyIn, 6*bodyIndex+4), ldS(c, c.FbodyIn, 6*bodyIndex+5));
            F = F - Fa;
        }
    }
    for (int k = 0; k < bc.nchild; ++k) {
        const BodyConst& cb = c.bodies[c.children[bc.childStart + k]];
        const CacheRef ch = cacheOf(c, cb.cacheBase);
        F = F + phi(ch.ld3(F_L), ch.ldSV(F_ZPLUS));
    }
    me.stSV(F_ZPLUS, F);
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const SV Hj = me.ldSV(F_H + 6*j);
        double tau = dot(Hj.w, F.w) + dot(Hj.v, F.v);
        if constexpr (WITH_VEL) { if (c.fmobIn) tau -= ldS(c, c.fmobIn, bc.u0 + j); }
        stS(c, c.vecOut, bc.u0 + j, tau);
    }
}

//==============================================================================================
// Joint-type dispatch.  The branch is uniform across a warp in the thread-per-instance plan
// (all lanes process the same body of different instances).
//==============================================================================================
#define SBK_DISPATCH_JOINT(jt, CALL)                                                \
    switch (jt) {                                                                   \
        case JT_PIN:       { constexpr int JT = JT_PIN;       CALL; } break;        \
        case JT_SLIDER:    { constexpr int JT = JT_SLIDER;    CALL; } break;        \
        case JT_UNIVERSAL: { constexpr int JT = JT_UNIVERSAL; CALL; } break;        \
        case JT_BALL:      { constexpr int JT = JT_BALL;      CALL; } break;        \
        case JT_FREE:      { constexpr int JT = JT_FREE;      CALL; } break;        \
        default: break;                                                             \
    }

SBK_HD void kinDispatch(const Ctx& c, int b) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (kinBody<JT>(c, bc)));
}
template <int MODE> SBK_HD void inwardDispatch(const Ctx& c, int b) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (inwardBody<JT, MODE>(c, bc, b)));
}
template <bool WITH_COR> SBK_HD void outwardDispatch(const Ctx& c, int b, double* udotDst, double* qddDst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (outwardBody<JT, WITH_COR>(c, bc, udotDst, qddDst)));
}
template <bool WITH_VEL> SBK_HD void idOutDispatch(const Ctx& c, int b) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (idOutBody<JT, WITH_VEL>(c, bc)));
}
template <bool WITH_VEL> SBK_HD void idInDispatch(const Ctx& c, int b) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (idInBody<JT, WITH_VEL>(c, bc, b)));
}

//==============================================================================================
// Per-instance drivers for the thread-per-instance plan: body index order is a valid
// base->tip order because a parent's MobilizedBodyIndex is always smaller than its child's.
//==============================================================================================
SBK_HD void tpiKinematics(const Ctx& c) { for (int b = 1; b < c.nb; ++b) kinDispatch(c, b); }
template <int MODE> SBK_HD void tpiInward(const Ctx& c) { for (int b = c.nb - 1; b >= 1; --b) inwardDispatch<MODE>(c, b); }
template <bool WITH_COR> SBK_HD void tpiOutward(const Ctx& c, double* udotDst, double* qddDst) {
    for (int b = 1; b < c.nb; ++b) outwardDispatch<WITH_COR>(c, b, udotDst, qddDst);
}
// One full derivative evaluation = System::realize(Acceleration) for the lowered system.
SBK_HD void tpiEvalDerivatives(const Ctx& c) {
    tpiKinematics(c);
    tpiInward<IN_ABI | IN_Z | IN_BIAS | IN_FORCES>(c);
    tpiOutward<true>(c, c.udot, c.qdotdot);
}

} // namespace sbkd
