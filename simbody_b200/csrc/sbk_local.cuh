// sbk_local.cuh -- body-frame ("local") articulated-body forward dynamics for the integrator path.
//
// What it computes.  The same qdot, udot as the reference's realize(Acceleration)
// (RigidBodyNodeSpec.cpp:249-446 + RigidBodyNodeSpec.h:229-333) for Pin / Slider / Universal / Ball /
// Free trees under Force::Gravity and mobility springs / dampers -- but every spatial quantity of a
// body is expressed in that body's own outboard frame M (origin Mo) instead of Ground.  The O(n)
// recursion is the reference's (sweep C + D inward, sweep E outward; abiCore / zCore are the very
// same functions with H replaced by the joint's motion subspace S expressed in M); what changes is
// the coordinate frame, and with it the work:
//   * no position kinematics in Ground at all: the parent->child transform is
//     X_{Mp -> M} = X_T * X_FM(q) with the CONSTANT X_T = X_MB(parent) * X_PF, so one sincos and one
//     3x3 product per body replace X_GB, H_PB_G, the re-expressed mass properties, Phi (sweep A);
//   * the rigid-body inertia about Mo in M is a constant of the body (no Rotation::reexpressSymMat33);
//   * S is a unit axis (Pin, Slider), the identity on the angular block (Ball) or the 6x6 identity
//     (Free) for the quasi-speeds u' = ~R_FM u: U = P*S is a COLUMN of the articulated inertia, D a
//     diagonal entry / block, so the reference's P*H, ~H*P*H, P*H*DI products disappear; a Free body
//     costs one 6x6 LDL^T solve (reference: getrf/getri + three 6x6 products);
//   * gravity enters as the base acceleration a_0 = -g (exact; no per-body force);
//   * body velocities are not stored between the sweeps: the inward sweep recovers the parent's
//     velocity from the child's by inverting the recurrence (6 + 18 flops), so per body and
//     evaluation only sin/cos (<= 4), G = U*DI (6 dof) and nu = DI*eps (dof) pass through HBM.
// udot is frame independent, so results agree with the ground-frame sweeps (and the reference) to
// rounding (measured <= 1e-13 relative; tests/test_oracle.py, tests/test_gpu_parity.py).
//
// Only the integrator uses this path; every getter / operator keeps the ground-frame FULL records.
#pragma once
#include <cstring>
#include "sbk_sweeps.cuh"

namespace sbkd {

// Per-body constants of the local path (batch-shared, 240 bytes; staged into shared memory).
struct LBody {
    double RT[9];           // R_T = R_MB(parent) * R_PF : rotation F(child's inboard frame) -> M(parent), row-major
    double pT[3];           // p_T = p_MB(parent) + R_MB(parent) * p_PF : origin of F in M(parent)
    double m;               // mass
    double h[3];            // m * (mass centre from Mo, in M)
    double I[6];            // inertia about Mo in M: xx yy zz xy xz yz
    int joint, parent, q0, u0;
    int flags, nchild, childStart, nforce;
    int forceStart, rec, parentLink, pad_;  // rec: first row of this body's scratch record; parentLink: first of the parent's V | A rows
    int childIA[4];         // first row of the P+ / z+ hand-over rows (IA) of the first four children (-1: none); further children through the tables
};
enum { BF_NO_RT = 32 };     // R_T is exactly the identity

// Scratch record rows of one body (CTA-blocked like the FULL records: [block][row][lane]):
//   SC 4 (sin/cos of the joint angles) | G 6d | NU d | V 6 | A 6 | V' 6 | IA 27 (P+ and z+ in the PARENT's frame)
// V / A are written only at branch points (a child that is not the next body reads them) and V at
// chain tips (the inward sweep starts its reverse velocity recurrence there); IA only by bodies
// whose parent is not the previous body.  V' is the second velocity buffer of the fused integrator
// (sbk_lrkm.cuh): its outward sweep reads the velocities of the state being evaluated from one buffer
// while it writes those of the next stage's state into the other.
enum { LR_SC = 0, LR_G = 4 };
SBK_HD constexpr int lrNU(int d)   { return 4 + 6*d; }
SBK_HD constexpr int lrV(int d)    { return 4 + 7*d; }
SBK_HD constexpr int lrA(int d)    { return 10 + 7*d; }
SBK_HD constexpr int lrIA(int d)   { return 22 + 7*d; }
SBK_HD constexpr int lrSize(int d) { return 49 + 7*d; }
enum { LR_VBUF = 12 };   // row distance between the two velocity buffers

// ---- spatial transforms between a parent's and a child's frame ------------------------------
// R: child -> parent rotation, p: child origin in parent coordinates.
SBK_HD SV xMotion(const M3& R, const V3 p, const SV a) {          // parent coords -> child coords
    SV r; r.w = mulT(R, a.w); r.v = mulT(R, a.v + cross(a.w, p)); return r;
}
SBK_HD SV xMotionInv(const M3& R, const V3 p, const SV a) {       // child coords -> parent coords
    SV r; r.w = mul(R, a.w); r.v = mul(R, a.v) - cross(r.w, p); return r;
}
SBK_HD SV xForce(const M3& R, const V3 p, const SV f) {           // child coords -> parent coords
    SV r; r.v = mul(R, f.v); r.w = mul(R, f.w) + cross(p, r.v); return r;
}
SBK_HD S3 rotSym(const M3& R, const S3& S) {                      // R S ~R
    const V3 t0 = mul(S, mk(R.a[0], R.a[1], R.a[2]));            // S * (row 0 of R)
    const V3 t1 = mul(S, mk(R.a[3], R.a[4], R.a[5]));
    const V3 t2 = mul(S, mk(R.a[6], R.a[7], R.a[8]));
    const V3 r0 = mk(R.a[0], R.a[1], R.a[2]), r1 = mk(R.a[3], R.a[4], R.a[5]), r2 = mk(R.a[6], R.a[7], R.a[8]);
    S3 o; o.xx = dot(r0, t0); o.yy = dot(r1, t1); o.zz = dot(r2, t2);
    o.xy = dot(r0, t1); o.xz = dot(r0, t2); o.yz = dot(r1, t2);
    return o;
}
SBK_HD ABI rotABI(const M3& R, const ABI& P) {                    // blockwise R (.) ~R
    ABI o; o.M = rotSym(R, P.M); o.J = rotSym(R, P.J); o.F = mulABt(mul(R, P.F), R); return o;
}
// spatial cross products (motion x motion, motion x* force)
SBK_HD SV crossMotion(const SV a, const SV b) { SV r; r.w = cross(a.w, b.w); r.v = cross(a.w, b.v) + cross(a.v, b.w); return r; }
SBK_HD SV crossForce(const SV a, const SV f)  { SV r; r.w = cross(a.w, f.w) + cross(a.v, f.v); r.v = cross(a.w, f.v); return r; }

// ---- joint maps in the child's M frame -------------------------------------------------------
template <int JT> struct LJoint {
    M3 R; V3 p;             // child (M) -> parent (M of the parent) rotation; Mo in parent coordinates
    M3 RFM;                 // Ball / Free: R_FM (quasi-speeds u' = ~R_FM u)
    double sb, cb;          // Universal: sin / cos of q1 (S column 0 = (cb, 0, sb))
};
enum { LSC_ROWS = 4 };
template <int JT> SBK_HD constexpr int lscCount() { return JT == JT_PIN ? 2 : JT == JT_UNIVERSAL ? 4 : 0; }
// G values a body keeps for the acceleration sweep: full columns (6 dof), Ball only Gv = ~F J^-1 (the angular block is I3), Free none (G = I6)
template <int JT> SBK_HD constexpr int lgCount() { return JT == JT_BALL ? 9 : JT == JT_FREE ? 0 : 6*JointDims<JT>::nu; }

// sc: sin/cos of the joint angles ((s, c) Pin; (s0, c0, s1, c1) Universal).  COMPUTE = true evaluates them
// from q and leaves them in sc (the velocity sweep stores them); false takes them from sc.
template <int JT, bool COMPUTE>
SBK_HD void ljoint(const LBody& bc, const double* q, double* sc, LJoint<JT>& k) {
    const bool noRT = (bc.flags & BF_NO_RT) != 0;
    const M3 RT = loadR(bc.RT); const V3 pT = mk(bc.pT[0], bc.pT[1], bc.pT[2]);
    k.p = pT;
    if constexpr (JT == JT_PIN) {                 // R_FM = Rz(q), RigidBodyNodeSpec_Pin.h:103-140
        if constexpr (COMPUTE) sincos(q[0], &sc[0], &sc[1]);
        const double s = sc[0], c = sc[1];
        if (noRT) { k.R.a[0] = c; k.R.a[1] = -s; k.R.a[2] = 0; k.R.a[3] = s; k.R.a[4] = c; k.R.a[5] = 0; k.R.a[6] = 0; k.R.a[7] = 0; k.R.a[8] = 1; }
        else {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                k.R.a[3*i]   = c*RT.a[3*i] + s*RT.a[3*i+1];
                k.R.a[3*i+1] = c*RT.a[3*i+1] - s*RT.a[3*i];
                k.R.a[3*i+2] = RT.a[3*i+2];
            }
        }
    } else if constexpr (JT == JT_SLIDER) {       // R_FM = I, p_FM = (q, 0, 0), RigidBodyNodeSpec_Slider.h:91-127
        k.R = RT; k.p = pT + q[0]*col(RT, 0);
    } else if constexpr (JT == JT_UNIVERSAL) {    // R_FM = Rx(q0) Ry(q1), RigidBodyNodeSpec_Universal.h:124-192
        if constexpr (COMPUTE) { sincos(q[0], &sc[0], &sc[1]); sincos(q[1], &sc[2], &sc[3]); }
        const double sa = sc[0], ca = sc[1], sb = sc[2], cb = sc[3];
        M3 F;
        F.a[0] = cb;     F.a[1] = 0;  F.a[2] = sb;
        F.a[3] = sb*sa;  F.a[4] = ca; F.a[5] = -sa*cb;
        F.a[6] = -sb*ca; F.a[7] = sa; F.a[8] = ca*cb;
        k.R = noRT ? F : mul(RT, F);
        k.sb = sb; k.cb = cb;
    } else {                                      // Ball / Free: RigidBodyNodeSpec_Ball.h:113-180, _Free.h:142-222
        const double oon = 1.0/sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
        k.RFM = rotFromQuat(q[0]*oon, q[1]*oon, q[2]*oon, q[3]*oon);
        k.R = noRT ? k.RFM : mul(RT, k.RFM);
        if constexpr (JT == JT_FREE) k.p = pT + (noRT ? mk(q[4], q[5], q[6]) : mul(RT, mk(q[4], q[5], q[6])));
    }
}

// Joint velocity v_J = S u' and velocity-product term c_J = (dS/dt) u' in M coordinates; up = quasi-speeds
// (u itself except for Ball / Free, where up = ~R_FM u blockwise).
template <int JT>
SBK_HD void ljointVel(const LJoint<JT>& k, const double* u, double* up, SV& vJ, SV& cJ) {
    vJ = zeroSV(); cJ = zeroSV();
    if constexpr (JT == JT_PIN)         { up[0] = u[0]; vJ.w.z = u[0]; }
    else if constexpr (JT == JT_SLIDER) { up[0] = u[0]; vJ.v.x = u[0]; }
    else if constexpr (JT == JT_UNIVERSAL) {      // S0 = ~R_FM x = (cb, 0, sb), S1 = y; dS0/dt = (-sb, 0, cb) u1
        up[0] = u[0]; up[1] = u[1];
        vJ.w = mk(k.cb*u[0], u[1], k.sb*u[0]);
        const double uu = u[0]*u[1];
        cJ.w = mk(-k.sb*uu, 0, k.cb*uu);
    } else {
        const V3 w = mulT(k.RFM, mk(u[0], u[1], u[2]));
        up[0] = w.x; up[1] = w.y; up[2] = w.z; vJ.w = w;
        if constexpr (JT == JT_FREE) {
            const V3 v = mulT(k.RFM, mk(u[3], u[4], u[5]));
            up[3] = v.x; up[4] = v.y; up[5] = v.z; vJ.v = v;
        }
    }
}
// Motion subspace columns in M coordinates (dense form for abiCore / zCore).
template <int JT>
SBK_HD void ljointS(const LJoint<JT>& k, SV* S) {
    constexpr int d = JointDims<JT>::nu;
#pragma unroll
    for (int j = 0; j < d; ++j) S[j] = zeroSV();
    if constexpr (JT == JT_PIN) S[0].w.z = 1;
    else if constexpr (JT == JT_SLIDER) S[0].v.x = 1;
    else if constexpr (JT == JT_UNIVERSAL) { S[0].w = mk(k.cb, 0, k.sb); S[1].w.y = 1; }
    else {
        S[0].w.x = 1; S[1].w.y = 1; S[2].w.z = 1;
        if constexpr (JT == JT_FREE) { S[3].v.x = 1; S[4].v.y = 1; S[5].v.z = 1; }
    }
}
// Applied mobility forces in quasi-speed coordinates: tau' = ~R_FM tau blockwise for Ball / Free.
template <int JT>
SBK_HD void ljointTau(const LJoint<JT>& k, double* f) {
    if constexpr (JT == JT_BALL || JT == JT_FREE) {
        const V3 a = mulT(k.RFM, mk(f[0], f[1], f[2])); f[0] = a.x; f[1] = a.y; f[2] = a.z;
        if constexpr (JT == JT_FREE) { const V3 b = mulT(k.RFM, mk(f[3], f[4], f[5])); f[3] = b.x; f[4] = b.y; f[5] = b.z; }
    }
}
// udot from the quasi-accelerations: u = R_FM u'  =>  udot_w = R_FM u'dot_w,  udot_v = R_FM (u'dot_v + w' x v')
template <int JT>
SBK_HD void ljointUdot(const LJoint<JT>& k, const double* up, const double* upd, double* udot) {
    constexpr int d = JointDims<JT>::nu;
    if constexpr (JT == JT_BALL || JT == JT_FREE) {
        const V3 a = mul(k.RFM, mk(upd[0], upd[1], upd[2])); udot[0] = a.x; udot[1] = a.y; udot[2] = a.z;
        if constexpr (JT == JT_FREE) {
            const V3 b = mul(k.RFM, mk(upd[3], upd[4], upd[5]) + cross(mk(up[0], up[1], up[2]), mk(up[3], up[4], up[5])));
            udot[3] = b.x; udot[4] = b.y; udot[5] = b.z;
        }
    } else {
#pragma unroll
        for (int i = 0; i < d; ++i) udot[i] = upd[i];
    }
}
// qdot = N(q) u (same as the ground-frame path: RigidBodyNodeSpec_Ball.h:187-212, _Free.h:226-259)
template <int JT>
SBK_HD void ljointQdot(const double* q, const double* u, double* qdot) {
    constexpr int d = JointDims<JT>::nu;
    if constexpr (JT == JT_BALL || JT == JT_FREE) {
        quatNTimes(q, mk(u[0], u[1], u[2]), qdot);
        if constexpr (JT == JT_FREE) { qdot[4] = u[3]; qdot[5] = u[4]; qdot[6] = u[5]; }
    } else {
#pragma unroll
        for (int i = 0; i < d; ++i) qdot[i] = u[i];
    }
}

// Rigid-body inertia about Mo in M as an articulated inertia (MassProperties.h:1256-1257 with constants)
SBK_HD ABI labiRigid(const LBody& bc) {
    ABI P;
    P.M.xx = bc.m; P.M.yy = bc.m; P.M.zz = bc.m; P.M.xy = 0; P.M.xz = 0; P.M.yz = 0;
    P.J.xx = bc.I[0]; P.J.yy = bc.I[1]; P.J.zz = bc.I[2]; P.J.xy = bc.I[3]; P.J.xz = bc.I[4]; P.J.yz = bc.I[5];
    P.F.a[0] = 0;        P.F.a[1] = -bc.h[2]; P.F.a[2] = bc.h[1];
    P.F.a[3] = bc.h[2];  P.F.a[4] = 0;        P.F.a[5] = -bc.h[0];
    P.F.a[6] = -bc.h[1]; P.F.a[7] = bc.h[0];  P.F.a[8] = 0;
    return P;
}
// Velocity-product bias force of the rigid body: v x* (I v)   (gyroscopic + centrifugal, RigidBodyNode.cpp:130-174 in M)
SBK_HD SV lbias(const LBody& bc, const SV v) {
    const V3 h = mk(bc.h[0], bc.h[1], bc.h[2]);
    S3 I; I.xx = bc.I[0]; I.yy = bc.I[1]; I.zz = bc.I[2]; I.xy = bc.I[3]; I.xz = bc.I[4]; I.yz = bc.I[5];
    SV Iv; Iv.w = mul(I, v.w) + cross(h, v.v); Iv.v = bc.m*v.v - cross(h, v.w);
    return crossForce(v, Iv);
}

// ---- scratch record access (same blocking as the FULL records) --------------------------------
template <bool BLK> SBK_HD CacheRefT<BLK> lrecOf(const Ctx& c, int inst, int row) {
    CacheRefT<BLK> r;
    r.p = c.cache + (BLK ? (long long)(inst >> 7)*c.cSpan + (long long)row*BLK_LANES + (inst & (BLK_LANES - 1))
                         : (long long)row*c.cStride + instOffset(c, inst));
    r.stride = c.cStride; return r;
}
// fcoef: per mobility u-slot j the lowered mobility forces as tau_j = A + B*q + C*u (springs, dampers, constant forces of the
// slot summed on the host: -k (q - q0) - c u + f = (k q0 + f) - k q - c u), 3 doubles per slot
struct LTables { const LBody* bodies; const int* children; const ForceConst* forces; const double* fcoef; };

// applied mobility forces of one body from the per-slot coefficients (cf. mobilityForces; Force.cpp:339-351,434-443)
template <int JT>
SBK_HD void lmobilityForces(const LBody& bc, const double* fcoef, const double* q, const double* u, double* f) {
    constexpr int d = JointDims<JT>::nu;
    constexpr bool QU = JT == JT_PIN || JT == JT_SLIDER || JT == JT_UNIVERSAL;     // springs exist only where q_j pairs with u_j
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double* k = fcoef + 3*(bc.u0 + j);
        f[j] = QU ? (k[0] + k[1]*q[j]) + k[2]*u[j] : k[0] + k[2]*u[j];
    }
}

// Carry column of one work item (shared memory on the device, [row][thread]):
//   outward sweeps : v (6) rows 0..5, a (6) rows 6..11 of the previous body
//   inward sweep   : P+ (21) z+ (6) of the next body in THIS body's frame rows 0..26, this body's velocity as
//                    recovered by that child rows 27..32
enum { LC_V = 0, LC_A = 6, LC_IA = 0, LC_VSELF = 27, LCARRY_ROWS = 33 };
SBK_HD void lcyStoreSV(double* cy, const SV a) {
    cy[0] = a.w.x; cy[1*SBK_CARRY_STRIDE] = a.w.y; cy[2*SBK_CARRY_STRIDE] = a.w.z;
    cy[3*SBK_CARRY_STRIDE] = a.v.x; cy[4*SBK_CARRY_STRIDE] = a.v.y; cy[5*SBK_CARRY_STRIDE] = a.v.z;
}
SBK_HD SV lcyLoadSV(const double* cy) {
    SV a; a.w = mk(cy[0], cy[1*SBK_CARRY_STRIDE], cy[2*SBK_CARRY_STRIDE]);
    a.v = mk(cy[3*SBK_CARRY_STRIDE], cy[4*SBK_CARRY_STRIDE], cy[5*SBK_CARRY_STRIDE]); return a;
}
SBK_HD void lcyStoreIA(double* cy, const ABI& P, const SV z) {
    const double v[27] = {P.M.xx, P.M.yy, P.M.zz, P.M.xy, P.M.xz, P.M.yz, P.J.xx, P.J.yy, P.J.zz, P.J.xy, P.J.xz, P.J.yz,
                          P.F.a[0], P.F.a[1], P.F.a[2], P.F.a[3], P.F.a[4], P.F.a[5], P.F.a[6], P.F.a[7], P.F.a[8],
                          z.w.x, z.w.y, z.w.z, z.v.x, z.v.y, z.v.z};
#pragma unroll
    for (int i = 0; i < 27; ++i) cy[i*SBK_CARRY_STRIDE] = v[i];
}
SBK_HD void lcyLoadIA(const double* cy, ABI& P, SV& z) {
    double v[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) v[i] = cy[i*SBK_CARRY_STRIDE];
    P.M.xx = v[0]; P.M.yy = v[1]; P.M.zz = v[2]; P.M.xy = v[3]; P.M.xz = v[4]; P.M.yz = v[5];
    P.J.xx = v[6]; P.J.yy = v[7]; P.J.zz = v[8]; P.J.xy = v[9]; P.J.xz = v[10]; P.J.yz = v[11];
#pragma unroll
    for (int i = 0; i < 9; ++i) P.F.a[i] = v[12+i];
    z.w = mk(v[21], v[22], v[23]); z.v = mk(v[24], v[25], v[26]);
}

template <int JT, bool BLK> SBK_HD void lloadCoords(const Ctx& c, const int inst, const LBody& bc, double* q, double* u) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
#pragma unroll
    for (int i = 0; i < NQ; ++i) q[i] = ldS<BLK>(c, inst, c.q, bc.q0 + i);
#pragma unroll
    for (int i = 0; i < d; ++i)  u[i] = ldS<BLK>(c, inst, c.u, bc.u0 + i);
}

//==============================================================================================
// Joint-specialised inward cores.  Same algebra as abiCore + zCore (RigidBodyNodeSpec.cpp:249-400) with the
// motion subspace S of the joint in M coordinates substituted symbolically:
//   z = P c + pA + sum z+(children);  eps = tau' - ~S z;  nu = DI eps;  G = P S DI;  z+ = z + G eps;  P+ = P - G ~(P S)
// Out: G / nu for the record (layout per joint, see lOutwardBody), and P+, z+ in the body's own frame.
//==============================================================================================
template <int d> struct LIn { double G[dim1(6*d)]; double nu[dim1(d)]; ABI PP; SV zP; bool ok; };

// P * c for a c whose z components are zero (Pin: c = v x (z u))
SBK_HD SV mulNoZ(const ABI& P, const double cwx, const double cwy, const double cvx, const double cvy) {
    SV r;
    r.w = mk(P.J.xx*cwx + P.J.xy*cwy + P.F.a[0]*cvx + P.F.a[1]*cvy,
             P.J.xy*cwx + P.J.yy*cwy + P.F.a[3]*cvx + P.F.a[4]*cvy,
             P.J.xz*cwx + P.J.yz*cwy + P.F.a[6]*cvx + P.F.a[7]*cvy);
    r.v = mk(P.F.a[0]*cwx + P.F.a[3]*cwy + P.M.xx*cvx + P.M.xy*cvy,
             P.F.a[1]*cwx + P.F.a[4]*cwy + P.M.xy*cvx + P.M.yy*cvy,
             P.F.a[2]*cwx + P.F.a[5]*cwy + P.M.xz*cvx + P.M.yz*cvy);
    return r;
}
// Pin: S = (z; 0).  U = P S = (J col z; F row z), D = Jzz.  P+ has a zero z row / column in J and a zero z row in F.
SBK_HD void lInPin(const ABI& P, const SV zsum, const SV pA, const SV v, const double u, const double tau, LIn<1>& o) {
    const double D = P.J.zz, DI = 1.0/D; o.ok = D != 0.0;
    const V3 Uw = mk(P.J.xz, P.J.yz, P.J.zz), Uv = mk(P.F.a[6], P.F.a[7], P.F.a[8]);
    const double gwx = Uw.x*DI, gwy = Uw.y*DI; const V3 gv = DI*Uv;
    o.G[0] = gwx; o.G[1] = gwy; o.G[2] = 1.0; o.G[3] = gv.x; o.G[4] = gv.y; o.G[5] = gv.z;
    const SV z = (mulNoZ(P, v.w.y*u, -(v.w.x*u), v.v.y*u, -(v.v.x*u)) + pA) + zsum;
    const double eps = tau - z.w.z;
    o.nu[0] = DI*eps;
    o.zP.w = mk(z.w.x + gwx*eps, z.w.y + gwy*eps, z.w.z + eps); o.zP.v = z.v + eps*gv;
    ABI& Q = o.PP;
    Q.J.xx = P.J.xx - gwx*Uw.x; Q.J.xy = P.J.xy - gwx*Uw.y; Q.J.yy = P.J.yy - gwy*Uw.y; Q.J.xz = 0; Q.J.yz = 0; Q.J.zz = 0;
    Q.F.a[0] = P.F.a[0] - gwx*Uv.x; Q.F.a[1] = P.F.a[1] - gwx*Uv.y; Q.F.a[2] = P.F.a[2] - gwx*Uv.z;
    Q.F.a[3] = P.F.a[3] - gwy*Uv.x; Q.F.a[4] = P.F.a[4] - gwy*Uv.y; Q.F.a[5] = P.F.a[5] - gwy*Uv.z;
    Q.F.a[6] = 0; Q.F.a[7] = 0; Q.F.a[8] = 0;
    Q.M.xx = P.M.xx - gv.x*Uv.x; Q.M.xy = P.M.xy - gv.x*Uv.y; Q.M.xz = P.M.xz - gv.x*Uv.z;
    Q.M.yy = P.M.yy - gv.y*Uv.y; Q.M.yz = P.M.yz - gv.y*Uv.z; Q.M.zz = P.M.zz - gv.z*Uv.z;
}
// Rotation of a Pin's P+ (zero z row / column in J, zero z row in F) by R, then the shift to the parent's origin.
SBK_HD ABI rotABIPin(const M3& R, const ABI& P) {
    ABI o; o.M = rotSym(R, P.M);
    double T[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {       // T = R(:,0:1) * J2
        T[i][0] = R.a[3*i]*P.J.xx + R.a[3*i+1]*P.J.xy;
        T[i][1] = R.a[3*i]*P.J.xy + R.a[3*i+1]*P.J.yy;
    }
    o.J.xx = T[0][0]*R.a[0] + T[0][1]*R.a[1]; o.J.xy = T[0][0]*R.a[3] + T[0][1]*R.a[4]; o.J.xz = T[0][0]*R.a[6] + T[0][1]*R.a[7];
    o.J.yy = T[1][0]*R.a[3] + T[1][1]*R.a[4]; o.J.yz = T[1][0]*R.a[6] + T[1][1]*R.a[7]; o.J.zz = T[2][0]*R.a[6] + T[2][1]*R.a[7];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[i][j] = R.a[3*i]*P.F.a[j] + R.a[3*i+1]*P.F.a[3+j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.F.a[3*i+j] = T[i][0]*R.a[3*j] + T[i][1]*R.a[3*j+1] + T[i][2]*R.a[3*j+2];
    return o;
}

// Inverse of a symmetric 3x3 by cofactors (the reference's 3x3 form, SmallMatrixMixed.h:938-958, on a symmetric input).
SBK_HD bool invSym3(const S3& A, S3& B) {
    const double c00 = A.yy*A.zz - A.yz*A.yz, c01 = A.xz*A.yz - A.xy*A.zz, c02 = A.xy*A.yz - A.xz*A.yy;
    const double det = A.xx*c00 + A.xy*c01 + A.xz*c02, ood = 1.0/det;
    B.xx = ood*c00; B.xy = ood*c01; B.xz = ood*c02;
    B.yy = ood*(A.xx*A.zz - A.xz*A.xz); B.yz = ood*(A.xy*A.xz - A.xx*A.yz); B.zz = ood*(A.xx*A.yy - A.xy*A.xy);
    return det != 0.0;
}
// Ball: S = (I3; 0) for the quasi-speeds w' = ~R_FM u.  D = J, G = (I3; ~F J^-1), P+ = (M - ~F J^-1 F) on the linear block only.
// o.G holds Gv = ~F J^-1 row-major (9 values).
SBK_HD void lInBall(const ABI& P, const SV zsum, const SV pA, const SV cc, const double* tau, LIn<3>& o) {
    S3 DI; o.ok = invSym3(P.J, DI);
    double Gv[9];                       // Gv[i][j] = sum_k F[k][i] DI[k][j]
    const double di[9] = {DI.xx, DI.xy, DI.xz, DI.xy, DI.yy, DI.yz, DI.xz, DI.yz, DI.zz};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Gv[3*i+j] = P.F.a[i]*di[j] + P.F.a[3+i]*di[3+j] + P.F.a[6+i]*di[6+j];
#pragma unroll
    for (int i = 0; i < 9; ++i) o.G[i] = Gv[i];
    const SV z = (mul(P, cc) + pA) + zsum;
    const V3 eps = mk(tau[0] - z.w.x, tau[1] - z.w.y, tau[2] - z.w.z);
    const V3 nu = mul(DI, eps);
    o.nu[0] = nu.x; o.nu[1] = nu.y; o.nu[2] = nu.z;
    o.zP.w = z.w + eps;
    o.zP.v = z.v + mk(Gv[0]*eps.x + Gv[1]*eps.y + Gv[2]*eps.z, Gv[3]*eps.x + Gv[4]*eps.y + Gv[5]*eps.z, Gv[6]*eps.x + Gv[7]*eps.y + Gv[8]*eps.z);
    // M+ = M - Gv F   (symmetric)
    S3& Mp = o.PP.M;
    Mp.xx = P.M.xx - (Gv[0]*P.F.a[0] + Gv[1]*P.F.a[3] + Gv[2]*P.F.a[6]);
    Mp.xy = P.M.xy - (Gv[0]*P.F.a[1] + Gv[1]*P.F.a[4] + Gv[2]*P.F.a[7]);
    Mp.xz = P.M.xz - (Gv[0]*P.F.a[2] + Gv[1]*P.F.a[5] + Gv[2]*P.F.a[8]);
    Mp.yy = P.M.yy - (Gv[3]*P.F.a[1] + Gv[4]*P.F.a[4] + Gv[5]*P.F.a[7]);
    Mp.yz = P.M.yz - (Gv[3]*P.F.a[2] + Gv[4]*P.F.a[5] + Gv[5]*P.F.a[8]);
    Mp.zz = P.M.zz - (Gv[6]*P.F.a[2] + Gv[7]*P.F.a[5] + Gv[8]*P.F.a[8]);
}
// An articulated inertia with a linear block only (what a Ball hands inward), rotated by R and shifted by s
// (ArticulatedInertia::shift semantics, MassProperties.cpp:101-127, with F = J = 0 on input).
SBK_HD ABI rotShiftMassOnly(const M3& R, const S3& M, const V3 s) {
    ABI o; o.M = rotSym(R, M);
    const V3 m0 = mk(o.M.xx, o.M.xy, o.M.xz), m1 = mk(o.M.xy, o.M.yy, o.M.yz), m2 = mk(o.M.xz, o.M.yz, o.M.zz);
    const V3 c0 = cross(s, m0), c1 = cross(s, m1), c2 = cross(s, m2);     // F' = s x M (column by column)
    o.F.a[0] = c0.x; o.F.a[1] = c1.x; o.F.a[2] = c2.x;
    o.F.a[3] = c0.y; o.F.a[4] = c1.y; o.F.a[5] = c2.y;
    o.F.a[6] = c0.z; o.F.a[7] = c1.z; o.F.a[8] = c2.z;
    #define G_(i,j) o.F.a[3*(i)+(j)]
    const double v0 = s.x, v1 = s.y, v2 = s.z;                             // J' = halfCrossDiff(s, 0, F')
    o.J.xx = v1*G_(0,2) - v2*G_(0,1);
    o.J.xy = v1*G_(1,2) - v2*G_(1,1);
    o.J.yy = v2*G_(1,0) - v0*G_(1,2);
    o.J.xz = v1*G_(2,2) - v2*G_(2,1);
    o.J.yz = v2*G_(2,0) - v0*G_(2,2);
    o.J.zz = v0*G_(2,1) - v1*G_(2,0);
    #undef G_
    return o;
}

// Free: S = I6 for the quasi-speeds (w', v') = (~R_FM w, ~R_FM v).  D = P, G = I6, P+ = 0, z+ = tau'.
// nu = P^-1 (tau' - z) by an unrolled LDL^T solve (P is symmetric positive definite); returns false if a pivot is not positive.
SBK_HD bool solveSym6(const ABI& P, const double* b, double* x) {
    double A[6][6];
    A[0][0] = P.J.xx; A[1][0] = P.J.xy; A[1][1] = P.J.yy; A[2][0] = P.J.xz; A[2][1] = P.J.yz; A[2][2] = P.J.zz;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[3+i][j] = P.F.a[3*j+i];          // lower-left block = ~F
    A[3][3] = P.M.xx; A[4][3] = P.M.xy; A[4][4] = P.M.yy; A[5][3] = P.M.xz; A[5][4] = P.M.yz; A[5][5] = P.M.zz;
    double L[6][6], W[6][6], od[6]; bool ok = true;                   // W = L * diag(d), od = 1/d
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double s = A[j][j];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < j) s -= L[j][k]*W[j][k];
        ok = ok && (s > 0.0); od[j] = 1.0/s;
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i > j) {
            double t = A[i][j];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < j) t -= W[i][k]*L[j][k];
            W[i][j] = t; L[i][j] = t*od[j];
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double t = b[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < i) t -= L[i][k]*y[k];
        y[i] = t;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double t = y[i]*od[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k > i) t -= L[k][i]*x[k];
        x[i] = t;
    }
    return ok;
}

// Universal: S0 = (cb, 0, sb; 0), S1 = (y; 0).
SBK_HD void lInUniversal(const ABI& P, const SV zsum, const SV pA, const SV cc, const double cb, const double sb, const double* tau, LIn<2>& o) {
    SV U0, U1;
    U0.w = mk(cb*P.J.xx + sb*P.J.xz, cb*P.J.xy + sb*P.J.yz, cb*P.J.xz + sb*P.J.zz);
    U0.v = mk(cb*P.F.a[0] + sb*P.F.a[6], cb*P.F.a[1] + sb*P.F.a[7], cb*P.F.a[2] + sb*P.F.a[8]);
    U1.w = mk(P.J.xy, P.J.yy, P.J.yz); U1.v = mk(P.F.a[3], P.F.a[4], P.F.a[5]);
    const double D00 = cb*U0.w.x + sb*U0.w.z, D01 = U0.w.y, D11 = P.J.yy;
    const double det = D00*D11 - D01*D01, ood = 1.0/det; o.ok = det != 0.0;
    const double I00 = ood*D11, I01 = -ood*D01, I11 = ood*D00;
    const SV G0 = I00*U0 + I01*U1, G1 = I01*U0 + I11*U1;
    const double g[12] = {G0.w.x, G0.w.y, G0.w.z, G0.v.x, G0.v.y, G0.v.z, G1.w.x, G1.w.y, G1.w.z, G1.v.x, G1.v.y, G1.v.z};
#pragma unroll
    for (int i = 0; i < 12; ++i) o.G[i] = g[i];
    const SV z = (mul(P, cc) + pA) + zsum;
    const double e0 = tau[0] - (cb*z.w.x + sb*z.w.z), e1 = tau[1] - z.w.y;
    o.nu[0] = I00*e0 + I01*e1; o.nu[1] = I01*e0 + I11*e1;
    o.zP = (z + e0*G0) + e1*G1;
    // P+ = P - G0 ~U0 - G1 ~U1
    const double gw[2][3] = {{G0.w.x, G0.w.y, G0.w.z}, {G1.w.x, G1.w.y, G1.w.z}}, gv[2][3] = {{G0.v.x, G0.v.y, G0.v.z}, {G1.v.x, G1.v.y, G1.v.z}};
    const double uw[2][3] = {{U0.w.x, U0.w.y, U0.w.z}, {U1.w.x, U1.w.y, U1.w.z}}, uv[2][3] = {{U0.v.x, U0.v.y, U0.v.z}, {U1.v.x, U1.v.y, U1.v.z}};
    ABI& Q = o.PP;
    #define SBK_UPD(i, j, A, B) ((A[0][i]*B[0][j]) + (A[1][i]*B[1][j]))
    Q.J.xx = P.J.xx - SBK_UPD(0, 0, gw, uw); Q.J.xy = P.J.xy - SBK_UPD(0, 1, gw, uw); Q.J.xz = P.J.xz - SBK_UPD(0, 2, gw, uw);
    Q.J.yy = P.J.yy - SBK_UPD(1, 1, gw, uw); Q.J.yz = P.J.yz - SBK_UPD(1, 2, gw, uw); Q.J.zz = P.J.zz - SBK_UPD(2, 2, gw, uw);
    Q.M.xx = P.M.xx - SBK_UPD(0, 0, gv, uv); Q.M.xy = P.M.xy - SBK_UPD(0, 1, gv, uv); Q.M.xz = P.M.xz - SBK_UPD(0, 2, gv, uv);
    Q.M.yy = P.M.yy - SBK_UPD(1, 1, gv, uv); Q.M.yz = P.M.yz - SBK_UPD(1, 2, gv, uv); Q.M.zz = P.M.zz - SBK_UPD(2, 2, gv, uv);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Q.F.a[3*i+j] = P.F.a[3*i+j] - SBK_UPD(i, j, gw, uv);
    #undef SBK_UPD
}

//==============================================================================================
// Prefetch of a body's inputs, one body step ahead.  A body step starts with loads whose values feed its
// first arithmetic (sin/cos rows -> rotation, coordinates, G / nu, the stage vectors): issued at the top of
// the step they cost a full trip to HBM with two warps per scheduler to cover it (ncu: 50% of the stall
// samples were long-scoreboard).  Every address is known from the body table alone, so the sweep drivers
// request the NEXT body's rows while the current one computes: asynchronous global->shared copies
// (cp.async, SASS LDGSTS) into one of two slots of the work item's shared-memory column, consumed with LDS.
// Layouts are compile-time per mobilizer kind (same functions on the producer and the consumer side).
//==============================================================================================
#ifndef SBK_LPF_ROWS
#define SBK_LPF_ROWS 32                     // a translation unit whose mobilizer kinds need fewer rows may shrink the slots
#endif
enum { LPF_ROWS = SBK_LPF_ROWS };           // rows per prefetch slot (two slots per work item)
SBK_HD void lpfCopy(double* dst, const double* src) {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    *dst = *src;
#endif
}
SBK_HD void lpfCommit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
SBK_HD void lpfWait() {                      // every request but the most recent one has landed
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
SBK_HD void lpfWaitAll() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// slot rows of the inward sweep: SC | q u | V (tips)
template <int JT> struct LPfIn  { enum { SC = 0, QU = lscCount<JT>(), V = QU + JointDims<JT>::nq + JointDims<JT>::nu }; };
// slot rows of the fused outward sweep: G | NU | SC | q u | Y slots | F0 slots | F2 or F3 slots; the heavy mobilizers leave
// the last one or two groups to direct loads (32 rows)
template <int JT> struct LPfOut {
    enum { NS = JointDims<JT>::nq + JointDims<JT>::nu, G = 0, NU = lgCount<JT>(), SC = NU + JointDims<JT>::nu, QU = SC + lscCount<JT>(),
           Y = QU + NS, F0 = Y + NS, F23 = F0 + NS,
           HAS_F0 = (JT != JT_BALL && JT != JT_FREE) ? 1 : 0, HAS_F23 = (JT == JT_PIN || JT == JT_SLIDER) ? 1 : 0 };
};
#define SBK_DISPATCH_LOCAL(JMASK, jt, CALL)                                                                         \
    switch (jt) {                                                                                                   \
        case JT_PIN:       if constexpr (((JMASK) & JM_PIN) != 0)       { constexpr int JT = JT_PIN;       CALL; } break; \
        case JT_SLIDER:    if constexpr (((JMASK) & JM_SLIDER) != 0)    { constexpr int JT = JT_SLIDER;    CALL; } break; \
        case JT_UNIVERSAL: if constexpr (((JMASK) & JM_UNIVERSAL) != 0) { constexpr int JT = JT_UNIVERSAL; CALL; } break; \
        case JT_BALL:      if constexpr (((JMASK) & JM_BALL) != 0)      { constexpr int JT = JT_BALL;      CALL; } break; \
        case JT_FREE:      if constexpr (((JMASK) & JM_FREE) != 0)      { constexpr int JT = JT_FREE;      CALL; } break; \
        default: break;                                                                                             \
    }

// The producer works on the NEXT body, whose kind is a run-time value: it dispatches once on the kind into straight-line copies
// with the group base addresses hoisted (one LDGSTS per row; a run-time row loop cost 19% of all executed instructions: ncu).
template <int N>
SBK_HD void lpfRowsN(double* dst, const double* src, const long long srcStride) {
#pragma unroll
    for (int i = 0; i < N; ++i) lpfCopy(dst + i*SBK_CARRY_STRIDE, src + i*srcStride);
}
template <int JT, bool BLK>
SBK_HD void lPrefetchInT(const Ctx& c, const LBody& bc, const int inst, double* pf, const double* S, const int vb) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    const long long rs = BLK ? BLK_LANES : me.stride, ss = BLK ? BLK_LANES : c.sStride;
    lpfRowsN<lscCount<JT>()>(pf + LPfIn<JT>::SC*SBK_CARRY_STRIDE, me.p + LR_SC*rs, rs);
    lpfRowsN<NQ>(pf + LPfIn<JT>::QU*SBK_CARRY_STRIDE, S + stateIndex<BLK>(c, inst, bc.q0), ss);
    lpfRowsN<d>(pf + (LPfIn<JT>::QU + NQ)*SBK_CARRY_STRIDE, S + stateIndex<BLK>(c, inst, c.nq + bc.u0), ss);
    if (bc.flags & BF_TIP) lpfRowsN<6>(pf + LPfIn<JT>::V*SBK_CARRY_STRIDE, me.p + (lrV(d) + vb)*rs, rs);
}
template <int JMASK, bool BLK>
SBK_HD void lPrefetchIn(const Ctx& c, const LBody& bc, const int inst, double* pf, const double* S, const int vb) {
    SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lPrefetchInT<JT, BLK>(c, bc, inst, pf, S, vb)));
}

//==============================================================================================
// Body steps (generic form: dense S through abiCore / zCore)
//==============================================================================================
// Velocity sweep, base -> tip: v = X v_parent + S u'.  Stores sin/cos; v only at tips / branch points.
template <int JT>
SBK_BODY void lVelBody(const Ctx& c, const LBody& bc, const int inst, double* cy, double* qdotDst, const int vb = 0) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    double q[dim1(NQ)], u[dim1(d)], up[dim1(d)], sc[LSC_ROWS];
    lloadCoords<JT, BLK>(c, inst, bc, q, u);
    LJoint<JT> k; ljoint<JT, true>(bc, q, sc, k);
#pragma unroll
    for (int i = 0; i < lscCount<JT>(); ++i) me.st(LR_SC + i, sc[i]);
    SV vP;
    if (bc.flags & BF_PARENT_PREV) vP = lcyLoadSV(cy + LC_V*SBK_CARRY_STRIDE);
    else vP = lrecOf<BLK>(c, inst, bc.parentLink).ldSV(vb);
    SV vJ, cJ; ljointVel<JT>(k, u, up, vJ, cJ);
    const SV v = xMotion(k.R, k.p, vP) + vJ;
    lcyStoreSV(cy + LC_V*SBK_CARRY_STRIDE, v);
    if (bc.flags & (BF_STORE_LINK | BF_TIP)) me.stSV(lrV(d) + vb, v);
    if (qdotDst) {
        double qdot[dim1(NQ)]; ljointQdot<JT>(q, u, qdot);
#pragma unroll
        for (int i = 0; i < NQ; ++i) stS<BLK>(c, inst, qdotDst, bc.q0 + i, qdot[i]);
    }
}

// Inward sweep, tip -> base: articulated inertia and bias force in M, G = U*DI and nu = DI*eps to the record,
// P+ / z+ handed to the parent in the parent's frame.
template <int JT>
SBK_BODY void lInwardBody(const Ctx& c, const LTables& T, const LBody& bc, const int inst, double* cy, const int vb = 0, const double* pf = nullptr) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    double q[dim1(NQ)], u[dim1(d)], up[dim1(d)], sc[LSC_ROWS];
    if (pf) {            // requested one body step ahead (lPrefetchIn)
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[i] = pf[(LPfIn<JT>::QU + i)*SBK_CARRY_STRIDE];
#pragma unroll
        for (int i = 0; i < d; ++i)  u[i] = pf[(LPfIn<JT>::QU + NQ + i)*SBK_CARRY_STRIDE];
#pragma unroll
        for (int i = 0; i < lscCount<JT>(); ++i) sc[i] = pf[(LPfIn<JT>::SC + i)*SBK_CARRY_STRIDE];
    } else {
        lloadCoords<JT, BLK>(c, inst, bc, q, u);
#pragma unroll
        for (int i = 0; i < lscCount<JT>(); ++i) sc[i] = me.ld(LR_SC + i);
    }
    LJoint<JT> k; ljoint<JT, false>(bc, q, sc, k);
    const bool haveCarryChild = !(bc.flags & BF_TIP);
    const SV v = haveCarryChild ? lcyLoadSV(cy + LC_VSELF*SBK_CARRY_STRIDE) : pf ? lcyLoadSV(pf + LPfIn<JT>::V*SBK_CARRY_STRIDE) : me.ldSV(lrV(d) + vb);
    SV vJ, cJ; ljointVel<JT>(k, u, up, vJ, cJ);
    const SV cc = cJ + crossMotion(v, vJ);

    ABI P = labiRigid(bc); SV z = zeroSV();
    if (haveCarryChild) { ABI cP; SV cz; lcyLoadIA(cy, cP, cz); addInto(P, cP); z = z + cz; }
    for (int j = haveCarryChild ? 1 : 0; j < bc.nchild; ++j) {
        int row;
        if (j < 4) row = bc.childIA[j];
        else { const LBody& cb = T.bodies[T.children[bc.childStart + j]]; row = cb.rec + lrIA(dofOfJoint(cb.joint)); }
        const CacheRefT<BLK> ch = lrecOf<BLK>(c, inst, row);
        addInto(P, ch.ldABI(0)); z = z + ch.ldSV(21);
    }
    const SV pA = lbias(bc, v);
    double f[dim1(d)];
    lmobilityForces<JT>(bc, T.fcoef, q, u, f);
    ljointTau<JT>(k, f);
    LIn<d> o; ABI Pp;
    if constexpr (JT == JT_PIN) {
        lInPin(P, z, pA, v, u[0], f[0], o);
        Pp = shiftABI(rotABIPin(k.R, o.PP), k.p);
    } else if constexpr (JT == JT_BALL) {
        lInBall(P, z, pA, cc, f, o);
        Pp = rotShiftMassOnly(k.R, o.PP.M, k.p);
    } else if constexpr (JT == JT_FREE) {
        const SV zz = (mul(P, cc) + pA) + z;
        const double eps[6] = {f[0] - zz.w.x, f[1] - zz.w.y, f[2] - zz.w.z, f[3] - zz.v.x, f[4] - zz.v.y, f[5] - zz.v.z};
        o.ok = solveSym6(P, eps, o.nu);
        o.zP.w = mk(f[0], f[1], f[2]); o.zP.v = mk(f[3], f[4], f[5]);
        std::memset(&Pp, 0, sizeof Pp);
    } else if constexpr (JT == JT_UNIVERSAL) {
        lInUniversal(P, z, pA, cc, k.cb, k.sb, f, o);
        Pp = shiftABI(rotABI(k.R, o.PP), k.p);
    } else {                                       // Slider: dense form
        SV S[dim1(d)]; ljointS<JT>(k, S);
        AbiOut<d> ao; double eps[dim1(d)];
        abiCore<d>(P, S, cc, pA, ao); o.ok = ao.ok;
        zCore<d>(S, ao.G, ao.zb + z, f, eps, o.zP);
#pragma unroll
        for (int j = 0; j < d; ++j) { const double g[6] = {ao.G[j].w.x, ao.G[j].w.y, ao.G[j].w.z, ao.G[j].v.x, ao.G[j].v.y, ao.G[j].v.z};
#pragma unroll
            for (int i = 0; i < 6; ++i) o.G[6*j+i] = g[i]; }
#pragma unroll
        for (int i = 0; i < d; ++i) { double sum = 0;
#pragma unroll
            for (int j = 0; j < d; ++j) sum += ao.DI[d*i+j]*eps[j];
            o.nu[i] = sum; }
        Pp = shiftABI(rotABI(k.R, ao.PP), k.p);
    }
    if (!o.ok) setSingular(c, inst);
#pragma unroll
    for (int i = 0; i < lgCount<JT>(); ++i) me.st(LR_G + i, o.G[i]);
#pragma unroll
    for (int i = 0; i < d; ++i) me.st(lrNU(d) + i, o.nu[i]);
    // hand P+, z+ to the parent, in the parent's frame at the parent's origin
    const SV zp = xForce(k.R, k.p, o.zP);
    if (bc.flags & BF_PARENT_PREV) {
        lcyStoreIA(cy, Pp, zp);
        lcyStoreSV(cy + LC_VSELF*SBK_CARRY_STRIDE, xMotionInv(k.R, k.p, v - vJ));
    } else {
        const CacheRefT<BLK> ia = lrecOf<BLK>(c, inst, bc.rec + lrIA(d));
        ia.stABI(0, Pp); ia.stSV(21, zp);
    }
}

// Acceleration sweep, base -> tip: a' = X a_parent, udot' = nu - ~G a', a = a' + S udot' + c.
template <int JT>
SBK_BODY void lOutwardBody(const Ctx& c, const LBody& bc, const int inst, double* cy, double* udotDst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    const CacheRefT<BLK> me = lrecOf<BLK>(c, inst, bc.rec);
    double q[dim1(NQ)], u[dim1(d)], up[dim1(d)], sc[LSC_ROWS], nu[dim1(d)], upd[dim1(d)], udot[dim1(d)];
    double G[dim1(lgCount<JT>())];
#pragma unroll
    for (int i = 0; i < lgCount<JT>(); ++i) G[i] = me.ld(LR_G + i);
#pragma unroll
    for (int j = 0; j < d; ++j) nu[j] = me.ld(lrNU(d) + j);
    lloadCoords<JT, BLK>(c, inst, bc, q, u);
#pragma unroll
    for (int i = 0; i < lscCount<JT>(); ++i) sc[i] = me.ld(LR_SC + i);
    LJoint<JT> k; ljoint<JT, false>(bc, q, sc, k);
    SV vP, aP;
    if (bc.flags & BF_PARENT_PREV) { vP = lcyLoadSV(cy + LC_V*SBK_CARRY_STRIDE); aP = lcyLoadSV(cy + LC_A*SBK_CARRY_STRIDE); }
    else { const CacheRefT<BLK> pa = lrecOf<BLK>(c, inst, bc.parentLink); vP = pa.ldSV(0); aP = pa.ldSV(6); }
    SV vJ, cJ; ljointVel<JT>(k, u, up, vJ, cJ);
    const SV v = xMotion(k.R, k.p, vP) + vJ;
    const SV cc = cJ + crossMotion(v, vJ);
    const SV aPlus = xMotion(k.R, k.p, aP);
    SV Su = zeroSV();
    const double ap[6] = {aPlus.w.x, aPlus.w.y, aPlus.w.z, aPlus.v.x, aPlus.v.y, aPlus.v.z};
    if constexpr (JT == JT_FREE) {
#pragma unroll
        for (int i = 0; i < 6; ++i) upd[i] = nu[i] - ap[i];
        Su.w = mk(upd[0], upd[1], upd[2]); Su.v = mk(upd[3], upd[4], upd[5]);
    } else if constexpr (JT == JT_BALL) {
#pragma unroll
        for (int j = 0; j < 3; ++j) upd[j] = nu[j] - (ap[j] + (G[j]*ap[3] + G[3+j]*ap[4] + G[6+j]*ap[5]));
        Su.w = mk(upd[0], upd[1], upd[2]);
    } else {
        SV S[dim1(d)]; ljointS<JT>(k, S);
#pragma unroll
        for (int i = 0; i < d; ++i) {
            double s = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) s += G[6*i+r]*ap[r];
            upd[i] = nu[i] - s;
            Su = Su + upd[i]*S[i];
        }
    }
    const SV a = (aPlus + Su) + cc;
    lcyStoreSV(cy + LC_V*SBK_CARRY_STRIDE, v); lcyStoreSV(cy + LC_A*SBK_CARRY_STRIDE, a);
    if (bc.flags & BF_STORE_LINK) me.stSV(lrA(d), a);
    ljointUdot<JT>(k, up, upd, udot);
    if (udotDst) {
#pragma unroll
        for (int i = 0; i < d; ++i) stS<BLK>(c, inst, udotDst, bc.u0 + i, udot[i]);
    }
}

// Ground's "record" (rows lrV(0).., written once per sweep by the sweep drivers): v = 0, a = -g.
SBK_HD void lGroundLinks(const Ctx& c, const LTables& T, const int inst, double* cy, bool withA) {
    constexpr bool BLK = SBK_DEV_BLK;
    SV a0 = zeroSV(); a0.v = mk(-c.gx, -c.gy, -c.gz);
    lcyStoreSV(cy + LC_V*SBK_CARRY_STRIDE, zeroSV());
    if (withA) lcyStoreSV(cy + LC_A*SBK_CARRY_STRIDE, a0);
    if (T.bodies[0].flags & BF_STORE_LINK) {
        const CacheRefT<BLK> g = lrecOf<BLK>(c, inst, T.bodies[0].rec);
        g.stSV(lrV(0), zeroSV());
        if (withA) g.stSV(lrA(0), a0);
    }
}

// One derivative evaluation with the local path (three sweeps; see sbk_lrkm.cuh for the fused integrator).
template <int JMASK = JM_MOBILE5> SBK_HD void lEvalDerivatives(const Ctx& c, const LTables& T, const int inst, double* cy, double* qdotDst, double* udotDst) {
    lGroundLinks(c, T, inst, cy, false);
#pragma unroll 1
    for (int b = 1; b < c.nb; ++b) { const LBody& bc = T.bodies[b]; SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lVelBody<JT>(c, bc, inst, cy, qdotDst))); }
#pragma unroll 1
    for (int b = c.nb - 1; b >= 1; --b) { const LBody& bc = T.bodies[b]; SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lInwardBody<JT>(c, T, bc, inst, cy))); }
    lGroundLinks(c, T, inst, cy, true);
#pragma unroll 1
    for (int b = 1; b < c.nb; ++b) { const LBody& bc = T.bodies[b]; SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lOutwardBody<JT>(c, bc, inst, cy, udotDst))); }
}

} // namespace sbkd
