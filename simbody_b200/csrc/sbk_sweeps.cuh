// sbk_sweeps.cuh -- per-body steps of the O(n) sweeps.
//
// Structure
//   *Core functions*  pure register arithmetic for ONE body of ONE instance (no memory
//                     traffic): kinCore (sweeps A+B), abiCore (sweep C + bias), zCore (sweep D),
//                     accCore (sweep E), qddCore, forceCore.
//   *Body wrappers*   FULL: move a body's data between the cores and the per-body cache records
//                     in HBM (API realize path of every plan; getters and operators read them).
//                     LEAN (integrator path of the thread-per-instance plan): reversible
//                     kinematics -- parent/child links ride in a per-thread carry column in shared
//                     memory, the inward sweep recovers each parent transform by inverting the
//                     recurrence, and only G and nu = DI*eps (7*dof doubles) per body reach HBM.
//   The register-resident fused plan (sbk_fused.cuh) calls the cores directly.
//
// Reference file:line each core replaces
//   kinCore   RigidBodyNodeSpec.h:229-333,554-569; RigidBodyNodeSpec.cpp:44-129;
//             RigidBodyNode.cpp:54-174; RigidBodyNodeSpec_{Pin,Slider,Universal,Ball,Free}.h;
//             RigidBodyNode_Weld.cpp:369-450 (Weld)
//   abiCore   RigidBodyNodeSpec.cpp:249-325 (ABI), RigidBodyNode.cpp:201-213 (P*a+b)
//   zCore     RigidBodyNodeSpec.cpp:355-400 (forward dynamics), :483-515 (M^-1 pass 1)
//   accCore   RigidBodyNodeSpec.cpp:408-446, :521-552
//   forces    Force_Gravity.cpp:514-569, Force.cpp:339-351,434-443
//   idOut/idIn (M*v, inverse dynamics) RigidBodyNodeSpec.cpp:566-695
//
// Per-body cache record (doubles), dof = d:
//   XGB 12 | VGB 6 | L 3 | MK 9 (c_G, G_G) | ACOR 6 | GYRO 6 | ZB 6 | PPLUS 21 | ZPLUS 6 | AGB 6 |
//   H 6d | G 6d | DI d*d | EPS d
#pragma once
#include "sbk_math.cuh"

// Body wrappers are force-inlined into their sweep drivers.  (An ABI call makes the callee save and
// restore every callee-saved register it touches -- ~100 registers = ~800 bytes of local-memory
// traffic per call for these FP64-heavy bodies, measured with ncu -- so the integrator has ONE inlined
// copy of the sweeps inside its stage loop, see tpiRkmStep.)
#define SBK_BODY SBK_HD

namespace sbkd {

enum { JT_GROUND = 0, JT_PIN = 1, JT_SLIDER = 2, JT_UNIVERSAL = 3, JT_BALL = 4, JT_FREE = 5, JT_WELD = 6,
       JT_TRANSLATION = 7, JT_CYLINDER = 8, JT_PLANAR = 9, JT_GIMBAL = 10,
       JT_BALL_EULER = 11, JT_FREE_EULER = 12 };   // internal: Ball / Free in Euler-angle mode (setUseEulerAngles)
enum { FK_SPRING = 2, FK_DAMPER = 3, FK_CONSTANT = 6 };

template <int JT> struct JointDims;
template <> struct JointDims<JT_PIN>       { enum { nq = 1, nu = 1 }; };
template <> struct JointDims<JT_SLIDER>    { enum { nq = 1, nu = 1 }; };
template <> struct JointDims<JT_UNIVERSAL> { enum { nq = 2, nu = 2 }; };
template <> struct JointDims<JT_BALL>      { enum { nq = 4, nu = 3 }; };
template <> struct JointDims<JT_FREE>      { enum { nq = 7, nu = 6 }; };
template <> struct JointDims<JT_WELD>      { enum { nq = 0, nu = 0 }; };   // RigidBodyNode_Weld.cpp:369: no q, no u
template <> struct JointDims<JT_TRANSLATION> { enum { nq = 3, nu = 3 }; };
template <> struct JointDims<JT_CYLINDER>  { enum { nq = 2, nu = 2 }; };
template <> struct JointDims<JT_PLANAR>    { enum { nq = 3, nu = 3 }; };
template <> struct JointDims<JT_GIMBAL>    { enum { nq = 3, nu = 3 }; };   // body-fixed x-y-z angles, u = qdot
template <> struct JointDims<JT_BALL_EULER> { enum { nq = 4, nu = 3 }; };  // 3 angles + 1 unused slot, u = w_FM
template <> struct JointDims<JT_FREE_EULER> { enum { nq = 7, nu = 6 }; };  // 3 angles, 3 translations, 1 unused slot
SBK_HD constexpr bool isBallLike(int jt) { return jt == JT_BALL || jt == JT_BALL_EULER; }
SBK_HD constexpr bool isFreeLike(int jt) { return jt == JT_FREE || jt == JT_FREE_EULER; }
SBK_HD constexpr bool isEulerKind(int jt) { return jt == JT_BALL_EULER || jt == JT_FREE_EULER; }
// array extent for a dof count that may be zero
SBK_HD constexpr int dim1(int d) { return d > 0 ? d : 1; }

// mobilizer-kind masks (bit JT_x set = kind present), see SBK_DISPATCH_JOINT_M
enum { JM_PIN = 1 << JT_PIN, JM_SLIDER = 1 << JT_SLIDER, JM_UNIVERSAL = 1 << JT_UNIVERSAL, JM_BALL = 1 << JT_BALL, JM_FREE = 1 << JT_FREE,
       JM_WELD = 1 << JT_WELD, JM_TRANSLATION = 1 << JT_TRANSLATION, JM_CYLINDER = 1 << JT_CYLINDER, JM_PLANAR = 1 << JT_PLANAR, JM_GIMBAL = 1 << JT_GIMBAL,
       JM_BALL_EULER = 1 << JT_BALL_EULER, JM_FREE_EULER = 1 << JT_FREE_EULER,
       JM_LIGHT = JM_PIN | JM_SLIDER | JM_UNIVERSAL | JM_WELD,                   // dof <= 2
       JM_MOBILE5 = JM_PIN | JM_SLIDER | JM_UNIVERSAL | JM_BALL | JM_FREE,       // the north_star mobilizer set
       JM_ALL = JM_MOBILE5 | JM_WELD | JM_TRANSLATION | JM_CYLINDER | JM_PLANAR | JM_GIMBAL | JM_BALL_EULER | JM_FREE_EULER,
       JM_LOCAL = 1 << 30 };   // integrator kernels only: body-frame sweeps (sbk_local.cuh) instead of the ground-frame ones

// cache record layout
enum { F_XGB = 0, F_VGB = 12, F_L = 18, F_MK = 21, F_ACOR = 30, F_GYRO = 36, F_ZB = 42,
       F_PPLUS = 48, F_ZPLUS = 69, F_AGB = 75, F_H = 81 };
SBK_HD constexpr int fG(int d)   { return F_H + 6*d; }
SBK_HD constexpr int fDI(int d)  { return F_H + 12*d; }
SBK_HD constexpr int fEPS(int d) { return F_H + 12*d + d*d; }
SBK_HD constexpr int cacheRecordSize(int d) { return F_H + 13*d + d*d; }
enum { CACHE_RECORD_MAX = F_H + 13*6 + 36 };   // 195

// body flags
enum { BF_PARENT_PREV = 1,   // parent is the body processed just before (index - 1): links ride in the carry
       BF_STORE_LINK  = 2,   // some child is NOT index + 1: links must also be stored in the cache
       BF_NO_R_PF     = 4,   // R_PF is exactly the identity (the reference's noR_PF template flag,
       BF_NO_R_MB     = 8,   // RigidBodyNodeSpec_Derived.cpp:50-64); likewise R_MB: skip the products
       BF_TIP         = 16 };// body index + 1 is NOT a child: the inward sweep of the integrator path starts its
                             // reverse kinematics here, from the X_GB, V_GB the outward sweep stored

// Per-body constants (batch-shared; staged into shared memory by the kernels).
struct BodyConst {
    double X_PF[12];        // R_PF row-major, p_PF
    double X_MB[12];        // ~X_BM (RigidBodyNode.h:1012)
    double mass;
    double com_B[3];
    double G_B[6];          // unit inertia about Bo in B: xx yy zz xy xz yz
    double p_BM[4];         // origin of the outboard frame M in B (X_BM.p; reaction forces are reported there); [3] pads
    long long cacheBase;    // element offset of this body's cache record (plan dependent)
    long long parentCacheBase;
    int joint, parent, q0, u0;
    int nchild, childStart, quat, nforce;
    int forceStart, level, flags, pad1;
};
struct ForceConst { int kind; int coord; double a; double b; };
// Force::TwoPointLinearSpring / TwoPointLinearDamper (Force.cpp:103-221): a force along the line between a station on body1 and one on body2
enum { TP_SPRING = 7, TP_DAMPER = 8 };
struct TwoPointConst { int kind, body1, body2, pad_; double a, b; double s1[3], s2[3]; long long cacheBase1, cacheBase2; };

// Context shared by every work item of a CTA (the device keeps ONE copy in shared memory, read
// with LDS: nothing per-thread lives in local memory).  SoA arrays are addressed
//   state : a[slot*sStride + inst*sInstStride]          cache : cache[base_b + k*cStride + inst*cInstStride]
struct Ctx {
    const BodyConst*  bodies;
    const int*        children;
    const ForceConst* forces;
    int nb, nq, nu, nquat;
    double gx, gy, gz;              // gravity vector g*d (Force_Gravity.cpp:532), 0 if none
    double*   cache;
    long long cStride, cInstStride, cSpan;   // cache offset of an instance: (inst >> cShift)*cSpan + (inst & cMask)*cInstStride
    int cShift, cMask;
    long long sStride, sInstStride, sSpan;   // state vectors: [slot][N] (sStride = N) or CTA-blocked (sSpan = rows*128 per block)
    const double* q; const double* u;
    double* qdot; double* udot; double* qdotdot; double* qerr;   // realize-path destinations (nullable)
    // operator inputs / outputs (nullable)
    const double* fmobIn;  const double* FbodyIn;   // applied forces for operator forms
    double* fmobOut; double* FbodyOut;              // force-subsystem results (getter)
    const double* vecIn; double* vecOut;            // generic nu-vectors for M, M^-1, residual
    int* status;                                    // per-instance status words [N]
    // two-point force elements (FULL records only): their body forces accumulate in f2 [nb*6][N] between the kinematics and the inward sweep
    const TwoPointConst* tp; int ntp; double* f2;
};

// The batch-shared tables as the integrator kernels see them: plain local pointers derived from
// the shared-memory staging buffer (or the global blob), so that after inlining the compiler knows
// their address space and emits LDS / LDG instead of generic loads.
struct Tables { const BodyConst* bodies; const int* children; const ForceConst* forces; };
SBK_HD Tables tablesOf(const Ctx& c) { Tables t; t.bodies = c.bodies; t.children = c.children; t.forces = c.forces; return t; }

// Links between consecutive bodies of one instance on the integrator (LEAN) path: a column of
// CARRY_ROWS doubles per work item in SHARED memory ([row][thread], conflict-free), reused by
// the three sweeps:
//   outward sweeps : X_GB (12) + V_GB (6) of the previous body, rows 0..17; A_GB (6) rows 18..23
//   inward sweep   : P+ (21) + z+ (6) + Phi.l (3) of the next body (a child), rows 0..29;
//                    X_GB (12) + V_GB (6) of THIS body as recovered by that child's reverse
//                    kinematics, rows 30..47
// Body b-1 is always the previous body of an outward sweep and b+1 of an inward one, so no
// bookkeeping is needed: BF_PARENT_PREV says whether the parent is b-1.
#ifndef SBK_GNU_ROWS
#define SBK_GNU_ROWS 14
#endif
enum { CY_A = 18, CY_SELF = 30, CY_PRE = 48, CY_GNU = 56, GNU_ROWS = SBK_GNU_ROWS, CY_LOOP = CY_GNU + 2*GNU_ROWS, CARRY_ROWS = CY_LOOP + 1 };
// row CY_LOOP: the body-loop counter of the sweep drivers (parked in shared memory across a body step instead of
// a register that the compiler would spill to local memory: its reload showed up as long-scoreboard stalls)
// rows 48..55: two coordinate preload slots; rows 56..83: two G / nu preload slots (acceleration sweep, dof <= 2)
#ifndef SBK_CARRY_STRIDE_DEVICE_THREADS
#define SBK_CARRY_STRIDE_DEVICE_THREADS 128      // threads per CTA of the kernels of this translation unit (sbk_ctree.cu: 256)
#endif
#if defined(__CUDA_ARCH__)
#define SBK_CARRY_STRIDE SBK_CARRY_STRIDE_DEVICE_THREADS
#else
#define SBK_CARRY_STRIDE 1
#endif
SBK_HD void cyStoreOut(double* cy, const M3& R, const V3 p, const SV V) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cy[i*SBK_CARRY_STRIDE] = R.a[i];
    cy[9*SBK_CARRY_STRIDE] = p.x; cy[10*SBK_CARRY_STRIDE] = p.y; cy[11*SBK_CARRY_STRIDE] = p.z;
    cy[12*SBK_CARRY_STRIDE] = V.w.x; cy[13*SBK_CARRY_STRIDE] = V.w.y; cy[14*SBK_CARRY_STRIDE] = V.w.z;
    cy[15*SBK_CARRY_STRIDE] = V.v.x; cy[16*SBK_CARRY_STRIDE] = V.v.y; cy[17*SBK_CARRY_STRIDE] = V.v.z;
}
SBK_HD void cyLoadOut(const double* cy, M3& R, V3& p, SV& V) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R.a[i] = cy[i*SBK_CARRY_STRIDE];
    p = mk(cy[9*SBK_CARRY_STRIDE], cy[10*SBK_CARRY_STRIDE], cy[11*SBK_CARRY_STRIDE]);
    V.w = mk(cy[12*SBK_CARRY_STRIDE], cy[13*SBK_CARRY_STRIDE], cy[14*SBK_CARRY_STRIDE]);
    V.v = mk(cy[15*SBK_CARRY_STRIDE], cy[16*SBK_CARRY_STRIDE], cy[17*SBK_CARRY_STRIDE]);
}
SBK_HD void cyStoreA(double* cy, const SV A) {
    cy[0] = A.w.x; cy[1*SBK_CARRY_STRIDE] = A.w.y; cy[2*SBK_CARRY_STRIDE] = A.w.z;
    cy[3*SBK_CARRY_STRIDE] = A.v.x; cy[4*SBK_CARRY_STRIDE] = A.v.y; cy[5*SBK_CARRY_STRIDE] = A.v.z;
}
SBK_HD SV cyLoadA(const double* cy) {
    SV A; A.w = mk(cy[0], cy[1*SBK_CARRY_STRIDE], cy[2*SBK_CARRY_STRIDE]);
    A.v = mk(cy[3*SBK_CARRY_STRIDE], cy[4*SBK_CARRY_STRIDE], cy[5*SBK_CARRY_STRIDE]); return A;
}
SBK_HD void cyStoreIn(double* cy, const ABI& P, const SV z, const V3 l) {
    const double v[30] = {P.M.xx, P.M.yy, P.M.zz, P.M.xy, P.M.xz, P.M.yz, P.J.xx, P.J.yy, P.J.zz, P.J.xy, P.J.xz, P.J.yz,
                          P.F.a[0], P.F.a[1], P.F.a[2], P.F.a[3], P.F.a[4], P.F.a[5], P.F.a[6], P.F.a[7], P.F.a[8],
                          z.w.x, z.w.y, z.w.z, z.v.x, z.v.y, z.v.z, l.x, l.y, l.z};
#pragma unroll
    for (int i = 0; i < 30; ++i) cy[i*SBK_CARRY_STRIDE] = v[i];
}
SBK_HD void cyLoadIn(const double* cy, ABI& P, SV& z, V3& l) {
    double v[30];
#pragma unroll
    for (int i = 0; i < 30; ++i) v[i] = cy[i*SBK_CARRY_STRIDE];
    P.M.xx = v[0]; P.M.yy = v[1]; P.M.zz = v[2]; P.M.xy = v[3]; P.M.xz = v[4]; P.M.yz = v[5];
    P.J.xx = v[6]; P.J.yy = v[7]; P.J.zz = v[8]; P.J.xy = v[9]; P.J.xz = v[10]; P.J.yz = v[11];
#pragma unroll
    for (int i = 0; i < 9; ++i) P.F.a[i] = v[12+i];
    z.w = mk(v[21], v[22], v[23]); z.v = mk(v[24], v[25], v[26]); l = mk(v[27], v[28], v[29]);
}

// Streaming global accesses bypass L1 (ld.global.cg / st.global.cg): the per-body records and the
// state vectors are touched once per sweep, while L1 is needed for the per-thread stack (carry
// links, Ctx, spills), which the streaming traffic would otherwise evict.
SBK_HD double gld(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
SBK_HD void gst(double* p, double v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}
// Thread-per-instance plans block their arrays by 128-instance CTA -- records [block][field][lane],
// integrator vectors [block][slot][lane] -- and the integrator path (BLK = true, device only) uses
// the row stride 128 as a COMPILE-TIME constant: every access of a body step is then
// `base + immediate`, with no per-access address arithmetic and no address registers.  (With a
// run-time stride each of the ~90 accesses of a body step needed its own 64-bit address; ncu
// showed the register pressure spilling freshly loaded values straight to local memory.)
enum { BLK_LANES = 128 };
#if defined(__CUDA_ARCH__)
#define SBK_DEV_BLK 1
#else
#define SBK_DEV_BLK 0
#endif
template <bool BLK> struct CacheRefT {     // accessor for one body's record
    double* p; long long stride;
    SBK_HD double ld(int k) const { return gld(p + (BLK ? (long long)k*BLK_LANES : (long long)k*stride)); }
    SBK_HD void   st(int k, double v) const { gst(p + (BLK ? (long long)k*BLK_LANES : (long long)k*stride), v); }
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
typedef CacheRefT<false> CacheRef;
// Instance offset into the cache: CTA-blocked records use cShift = 7 (block*cSpan + lane); the
// level-parallel plan cShift = 0 (inst*cSpan); the host emulation cShift = 30 (offset = inst).
SBK_HD long long instOffset(const Ctx& c, int inst) { return (long long)(inst >> c.cShift)*c.cSpan + (long long)(inst & c.cMask)*c.cInstStride; }
template <bool BLK> SBK_HD CacheRefT<BLK> cacheOf(const Ctx& c, int inst, long long base) {
    CacheRefT<BLK> r;
    r.p = c.cache + base + (BLK ? (long long)(inst >> 7)*c.cSpan + (inst & (BLK_LANES - 1)) : instOffset(c, inst));
    r.stride = c.cStride; return r;
}
// State / integrator vectors: BLK -> [block][slot][lane] with span sSpan per block; else [slot][N].
template <bool BLK> SBK_HD long long stateIndex(const Ctx& c, int inst, int slot) {
    return BLK ? (long long)(inst >> 7)*c.sSpan + (long long)slot*BLK_LANES + (inst & (BLK_LANES - 1))
               : (long long)slot*c.sStride + (long long)inst*c.sInstStride;
}
template <bool BLK> SBK_HD double ldS(const Ctx& c, int inst, const double* a, int slot) { return gld(a + stateIndex<BLK>(c, inst, slot)); }
template <bool BLK> SBK_HD void   stS(const Ctx& c, int inst, double* a, int slot, double v) { gst(a + stateIndex<BLK>(c, inst, slot), v); }

SBK_HD int dofOfJoint(int jt) { return isFreeLike(jt) ? 6 : (isBallLike(jt) || jt == JT_TRANSLATION || jt == JT_PLANAR || jt == JT_GIMBAL) ? 3 : (jt == JT_UNIVERSAL || jt == JT_CYLINDER) ? 2
                                      : (jt == JT_GROUND || jt == JT_WELD) ? 0 : 1; }
SBK_HD int nqOfJoint(int jt)  { return isFreeLike(jt) ? 7 : isBallLike(jt) ? 4 : (jt == JT_TRANSLATION || jt == JT_PLANAR || jt == JT_GIMBAL) ? 3 : (jt == JT_UNIVERSAL || jt == JT_CYLINDER) ? 2
                                      : (jt == JT_GROUND || jt == JT_WELD) ? 0 : 1; }
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
// Body-fixed x-y-z Euler angles: qdot = N_P(q) * w with sc = {c0, s0, c1, s1, 1/c1} (Rotation.h:395-406)
SBK_HD V3 eulerNTimes(const double* sc, V3 w) {
    const double c0 = sc[0], s0 = sc[1], s1 = sc[3], oocosy = sc[4];
    const double t = (s0*w.y - c0*w.z)*oocosy;
    return mk(w.x + t*s1, c0*w.y + s0*w.z, -t);
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
//                                        CORES
//==============================================================================================
template <int d> struct KinOut {     // results of sweeps A+B for one body
    M3 R; V3 p; SV V;                // X_GB, V_GB
    SV H[dim1(d)]; V3 l; V3 c; S3 G; // H_PB_G columns, Phi.l, com in G, unit inertia in G
    SV acor, gyro;                   // mobilizer coriolis acceleration a, gyroscopic force b
};

// The H_FM columns of the five mobilizers are unit frame axes (or zero) except Universal's second
// one; these helpers spell out products with them so that no flop is spent multiplying by a
// literal 0 or 1 (the compiler may not drop x*0: NaN/Inf semantics).  Same values as the dense
// expressions up to the sign of zeros.
template <int A> SBK_HD V3 axisCross(V3 v) {          // e_A x v
    if constexpr (A == 0) return mk(0, -v.z, v.y);
    else if constexpr (A == 1) return mk(v.z, 0, -v.x);
    else return mk(-v.y, v.x, 0);
}
template <int A> SBK_HD V3 mulSkip(const M3& R, V3 t) {   // R * t for a t whose component A is zero
    if constexpr (A == 0) return mk(R.a[1]*t.y + R.a[2]*t.z, R.a[4]*t.y + R.a[5]*t.z, R.a[7]*t.y + R.a[8]*t.z);
    else if constexpr (A == 1) return mk(R.a[0]*t.x + R.a[2]*t.z, R.a[3]*t.x + R.a[5]*t.z, R.a[6]*t.x + R.a[8]*t.z);
    else return mk(R.a[0]*t.x + R.a[1]*t.y, R.a[3]*t.x + R.a[4]*t.y, R.a[6]*t.x + R.a[7]*t.y);
}

// Position kinematics that depend on the mobilizer coordinates only (no parent quantities):
// X_FM, H_FM, and X_PB = X_PF * X_FM * X_MB (RigidBodyNodeSpec.h:554-569).
template <int d> struct KinLocal {
    M3 R_FM; V3 Hw[dim1(d)], Hv[dim1(d)];   // H_FM columns (angular, linear), expressed in F
    V3 r;                            // r_MB_F = R_FM * p_MB
    M3 R_PB; V3 p_PB;                // X_PB
    double qerr;                     // |q| - 1 for quaternion mobilizers
    double sc[5];                    // Gimbal / Euler-angle mode: c0, s0, c1, s1, 1/c1 (the reference's q pool)
};

template <int JT>
SBK_HD void kinLocal(const BodyConst& bc, const double* q, KinLocal<JointDims<JT>::nu>& k) {
    constexpr int d = JointDims<JT>::nu;
    V3 p_FM = zero3();
#pragma unroll
    for (int j = 0; j < d; ++j) { k.Hw[j] = zero3(); k.Hv[j] = zero3(); }
    k.qerr = 0;
    M3& R_FM = k.R_FM;

    if constexpr (JT == JT_PIN || JT == JT_CYLINDER || JT == JT_PLANAR) {   // R_FM = Rz(q0): _Pin.h:103-140, _Cylinder.h:110-139, _Planar.h:126-157
        double s, co; sincos(q[0], &s, &co);
        R_FM.a[0] = co; R_FM.a[1] = -s; R_FM.a[2] = 0;
        R_FM.a[3] = s;  R_FM.a[4] = co; R_FM.a[5] = 0;
        R_FM.a[6] = 0;  R_FM.a[7] = 0;  R_FM.a[8] = 1;
        k.Hw[0] = mk(0, 0, 1);
        if constexpr (JT == JT_CYLINDER) { p_FM = mk(0, 0, q[1]); k.Hv[1] = mk(0, 0, 1); }
        if constexpr (JT == JT_PLANAR)   { p_FM = mk(q[1], q[2], 0); k.Hv[1] = mk(1, 0, 0); k.Hv[2] = mk(0, 1, 0); }
    } else if constexpr (JT == JT_BALL_EULER || JT == JT_FREE_EULER) {   // _Ball.h:118-160, _Free.h:147-180: x-y-z angles, H_FM = I
        double s0, c0, s1, c1, s2, c2; sincos(q[0], &s0, &c0); sincos(q[1], &s1, &c1); sincos(q[2], &s2, &c2);
        const double s0s1 = s0*s1, s2c0 = s2*c0, c0c2 = c0*c2, nc1 = -c1;
        R_FM.a[0] = c1*c2;             R_FM.a[1] = s2*nc1;            R_FM.a[2] = s1;
        R_FM.a[3] = s2c0 + s0s1*c2;    R_FM.a[4] = c0c2 - s0s1*s2;    R_FM.a[5] = s0*nc1;
        R_FM.a[6] = s0*s2 - s1*c0c2;   R_FM.a[7] = s0*c2 + s1*s2c0;   R_FM.a[8] = c0*c1;
        k.Hw[0] = mk(1, 0, 0); k.Hw[1] = mk(0, 1, 0); k.Hw[2] = mk(0, 0, 1);
        k.sc[0] = c0; k.sc[1] = s0; k.sc[2] = c1; k.sc[3] = s1; k.sc[4] = 1/c1;      // trouble at +-90 degrees, as in the reference
        if constexpr (JT == JT_FREE_EULER) {
            p_FM = mk(q[3], q[4], q[5]);
            k.Hv[3] = mk(1, 0, 0); k.Hv[4] = mk(0, 1, 0); k.Hv[5] = mk(0, 0, 1);
        }
    } else if constexpr (JT == JT_GIMBAL) {       // RigidBodyNodeSpec_Gimbal.h:108-176, Rotation.h:342-349
        double s0, c0, s1, c1, s2, c2; sincos(q[0], &s0, &c0); sincos(q[1], &s1, &c1); sincos(q[2], &s2, &c2);
        const double s0s1 = s0*s1, s2c0 = s2*c0, c0c2 = c0*c2, nc1 = -c1;
        R_FM.a[0] = c1*c2;             R_FM.a[1] = s2*nc1;            R_FM.a[2] = s1;
        R_FM.a[3] = s2c0 + s0s1*c2;    R_FM.a[4] = c0c2 - s0s1*s2;    R_FM.a[5] = s0*nc1;
        R_FM.a[6] = s0*s2 - s1*c0c2;   R_FM.a[7] = s0*c2 + s1*s2c0;   R_FM.a[8] = c0*c1;
        k.Hw[0] = mk(1, 0, 0); k.Hw[1] = mk(0, c0, s0); k.Hw[2] = mk(s1, -s0*c1, c0*c1);
        k.sc[0] = c0; k.sc[1] = s0; k.sc[2] = c1; k.sc[3] = s1;
    } else if constexpr (JT == JT_TRANSLATION) {  // RigidBodyNodeSpec_Translation.h:100-130
        R_FM = identity3(); p_FM = mk(q[0], q[1], q[2]);
        k.Hv[0] = mk(1, 0, 0); k.Hv[1] = mk(0, 1, 0); k.Hv[2] = mk(0, 0, 1);
    } else if constexpr (JT == JT_WELD) {         // RigidBodyNode_Weld.cpp:390-398: X_FM = I
        R_FM = identity3();
    } else if constexpr (JT == JT_SLIDER) {       // RigidBodyNodeSpec_Slider.h:91-127
        R_FM = identity3(); p_FM = mk(q[0], 0, 0);
        k.Hv[0] = mk(1, 0, 0);
    } else if constexpr (JT == JT_UNIVERSAL) {    // RigidBodyNodeSpec_Universal.h:124-192, Rotation.cpp:241-264
        double s1, c1, s2, c2; sincos(q[0], &s1, &c1); sincos(q[1], &s2, &c2);
        R_FM.a[0] = c2;      R_FM.a[1] = 0;  R_FM.a[2] = s2;
        R_FM.a[3] = s2*s1;   R_FM.a[4] = c1; R_FM.a[5] = -s1*c2;
        R_FM.a[6] = -s2*c1;  R_FM.a[7] = s1; R_FM.a[8] = c1*c2;
        k.Hw[0] = mk(1, 0, 0);
        k.Hw[1] = col(R_FM, 1);
    } else {                                      // Ball / Free: RigidBodyNodeSpec_Ball.h:113-180, _Free.h:142-222
        const double quatLen = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
        k.qerr = quatLen - 1.0;
        const double oon = 1.0/quatLen;
        R_FM = rotFromQuat(q[0]*oon, q[1]*oon, q[2]*oon, q[3]*oon);
        k.Hw[0] = mk(1, 0, 0); k.Hw[1] = mk(0, 1, 0); k.Hw[2] = mk(0, 0, 1);
        if constexpr (JT == JT_FREE) {
            p_FM = mk(q[4], q[5], q[6]);
            k.Hv[3] = mk(1, 0, 0); k.Hv[4] = mk(0, 1, 0); k.Hv[5] = mk(0, 0, 1);
        }
    }

    // ---- calcBodyTransforms, parent-independent part (RigidBodyNodeSpec.h:554-569) ------------
    const M3 R_PF = loadR(bc.X_PF), R_MB = loadR(bc.X_MB);
    const V3 p_PF = loadP(bc.X_PF), p_MB = loadP(bc.X_MB);
    // Multiplying by an exact identity is skipped (same values up to the sign of zeros), as the
    // reference does through its noR_PF / noX_MB template flags.
    const bool noRPF = (bc.flags & BF_NO_R_PF) != 0, noRMB = (bc.flags & BF_NO_R_MB) != 0;
    M3 R_FB;
    if constexpr (JT == JT_PIN || JT == JT_CYLINDER || JT == JT_PLANAR) {   // R_FM = Rz(q0): products written out without the 0 / 1 entries
        const double co = R_FM.a[0], si = R_FM.a[3];
        k.r = mk(co*p_MB.x - si*p_MB.y, si*p_MB.x + co*p_MB.y, p_MB.z);
        if (noRMB) R_FB = R_FM;
        else {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                R_FB.a[j]     = co*R_MB.a[j] - si*R_MB.a[3+j];
                R_FB.a[3 + j] = si*R_MB.a[j] + co*R_MB.a[3+j];
                R_FB.a[6 + j] = R_MB.a[6+j];
            }
        }
    } else if constexpr (JT == JT_SLIDER || JT == JT_WELD || JT == JT_TRANSLATION) {   // R_FM = I
        k.r = p_MB; R_FB = R_MB;
    } else {
        k.r = mul(R_FM, p_MB);                         // r_MB_F = R_FM * p_MB
        R_FB = noRMB ? R_FM : mul(R_FM, R_MB);
    }
    const V3 p_FB = p_FM + k.r;
    k.R_PB = noRPF ? R_FB : mul(R_PF, R_FB);
    k.p_PB = p_PF + (noRPF ? p_FB : mul(R_PF, p_FB));
}

// H = R_GF (H_FM + H_MB_F), H_MB_F[1] = -r x H_FM[0]  (RigidBodyNodeSpec.cpp:44-74)
template <int JT>
SBK_HD void jointH(const M3& R_GF, const KinLocal<JointDims<JT>::nu>& k, SV* H) {
    const V3 r = k.r;
    #define SBK_ROTCOL(j, A) { H[j].w = col(R_GF, A); H[j].v = mulSkip<A>(R_GF, axisCross<A>(r)); }
    #define SBK_TRCOL(j, A)  { H[j].w = zero3(); H[j].v = col(R_GF, A); }
    if constexpr (JT == JT_WELD) { (void)r; (void)H; }
    else if constexpr (JT == JT_PIN) SBK_ROTCOL(0, 2)
    else if constexpr (JT == JT_SLIDER) SBK_TRCOL(0, 0)
    else if constexpr (JT == JT_TRANSLATION) { SBK_TRCOL(0, 0) SBK_TRCOL(1, 1) SBK_TRCOL(2, 2) }
    else if constexpr (JT == JT_CYLINDER) { SBK_ROTCOL(0, 2) SBK_TRCOL(1, 2) }
    else if constexpr (JT == JT_PLANAR) { SBK_ROTCOL(0, 2) SBK_TRCOL(1, 0) SBK_TRCOL(2, 1) }
    else if constexpr (JT == JT_UNIVERSAL) {
        SBK_ROTCOL(0, 0)
        H[1].w = mul(R_GF, k.Hw[1]); H[1].v = mul(R_GF, cross(k.Hw[1], r));
    } else if constexpr (JT == JT_GIMBAL) {
        SBK_ROTCOL(0, 0)
        H[1].w = mul(R_GF, k.Hw[1]); H[1].v = mul(R_GF, cross(k.Hw[1], r));
        H[2].w = mul(R_GF, k.Hw[2]); H[2].v = mul(R_GF, cross(k.Hw[2], r));
    } else {
        SBK_ROTCOL(0, 0) SBK_ROTCOL(1, 1) SBK_ROTCOL(2, 2)
        if constexpr (isFreeLike(JT)) { SBK_TRCOL(3, 0) SBK_TRCOL(4, 1) SBK_TRCOL(5, 2) }
    }
    #undef SBK_ROTCOL
    #undef SBK_TRCOL
}
// w_FM = H_FM(angular) u
template <int JT> SBK_HD V3 jointWFM(const KinLocal<JointDims<JT>::nu>& k, const double* u) {
    if constexpr (JT == JT_PIN || JT == JT_CYLINDER || JT == JT_PLANAR) return mk(0, 0, u[0]);
    else if constexpr (JT == JT_SLIDER || JT == JT_WELD || JT == JT_TRANSLATION) return zero3();
    else if constexpr (JT == JT_UNIVERSAL) return mk(u[0], 0, 0) + u[1]*k.Hw[1];
    else if constexpr (JT == JT_GIMBAL) return (mk(u[0], 0, 0) + u[1]*k.Hw[1]) + u[2]*k.Hw[2];
    else return mk(u[0], u[1], u[2]);
}

// Everything that needs the parent's X_GP, V_GP.
template <int JT>
SBK_HD void kinGlobal(const BodyConst& bc, const KinLocal<JointDims<JT>::nu>& k, const double* q, const double* u,
                      const M3& R_GP, const V3 p_GP, const SV V_GP,
                      KinOut<JointDims<JT>::nu>& o, double* qdot) {
    constexpr int d = JointDims<JT>::nu;
    const bool noRPF = (bc.flags & BF_NO_R_PF) != 0;
    const V3 r = k.r;

    o.R = mul(R_GP, k.R_PB);
    o.l = mul(R_GP, k.p_PB);                          // Phi: p_PB_G (RigidBodyNode.cpp:61)
    o.p = p_GP + o.l;

    // ---- H = R_GF (H_FM + H_MB_F)  (RigidBodyNodeSpec.cpp:44-74) -----------------------------
    const M3 R_GF = noRPF ? R_GP : mul(R_GP, loadR(bc.X_PF));
    jointH<JT>(R_GF, k, o.H);

    // ---- mass properties in Ground (RigidBodyNode.cpp:54-84) ---------------------------------
    S3 G_B; G_B.xx = bc.G_B[0]; G_B.yy = bc.G_B[1]; G_B.zz = bc.G_B[2]; G_B.xy = bc.G_B[3]; G_B.xz = bc.G_B[4]; G_B.yz = bc.G_B[5];
    o.G = reexpressSym(o.R, G_B);
    o.c = mul(o.R, mk(bc.com_B[0], bc.com_B[1], bc.com_B[2]));

    // ---- velocity (RigidBodyNodeSpec.h:305-333) ------------------------------------------------
    const V3 w_FM = jointWFM<JT>(k, u);
    SV V_PB = zeroSV();
    if constexpr (d > 0) V_PB = u[0]*o.H[0];
#pragma unroll
    for (int j = 1; j < d; ++j) V_PB = V_PB + u[j]*o.H[j];

    // HDot (RigidBodyNodeSpec.cpp:82-129) and VD = HDot*u.  HDot_FM is zero for every column but
    // Universal's second (_Universal.h:176-190), and translational columns have no angular part.
    const V3 w_GP = V_GP.w, v_GP = V_GP.v;
    const V3 wxr = cross(w_FM, r);
    SV VD = zeroSV();
    #define SBK_ROTCOL_D(j, A) { SV HD; HD.w = cross(w_GP, o.H[j].w); \
                                 HD.v = mulSkip<A>(R_GF, axisCross<A>(wxr)) + cross(w_GP, o.H[j].v); VD = VD + u[j]*HD; }
    #define SBK_TRCOL_D(j)     { SV HD; HD.w = zero3(); HD.v = cross(w_GP, o.H[j].v); VD = VD + u[j]*HD; }
    if constexpr (JT == JT_WELD) { (void)wxr; }
    else if constexpr (JT == JT_PIN) SBK_ROTCOL_D(0, 2)
    else if constexpr (JT == JT_SLIDER) SBK_TRCOL_D(0)
    else if constexpr (JT == JT_TRANSLATION) { SBK_TRCOL_D(0) SBK_TRCOL_D(1) SBK_TRCOL_D(2) }
    else if constexpr (JT == JT_CYLINDER) { SBK_ROTCOL_D(0, 2) SBK_TRCOL_D(1) }
    else if constexpr (JT == JT_PLANAR) { SBK_ROTCOL_D(0, 2) SBK_TRCOL_D(1) SBK_TRCOL_D(2) }
    else if constexpr (JT == JT_UNIVERSAL) {
        SBK_ROTCOL_D(0, 0)
        const V3 HDw1 = cross(w_FM, col(k.R_FM, 1));
        SV HD;
        HD.w = mul(R_GF, HDw1) + cross(w_GP, o.H[1].w);
        HD.v = mul(R_GF, cross(HDw1, r) + cross(k.Hw[1], wxr)) + cross(w_GP, o.H[1].v);
        VD = VD + u[1]*HD;
    } else if constexpr (JT == JT_GIMBAL) {          // _Gimbal.h:150-176 with qdot = u
        SBK_ROTCOL_D(0, 0)
        const double c0 = k.sc[0], s0 = k.sc[1], c1 = k.sc[2], s1 = k.sc[3];
        const double dc0 = -s0*u[0], dc1 = -s1*u[1], ds0 = c0*u[0], ds1 = c1*u[1];
        const V3 HDwj[2] = { mk(0, dc0, ds0), mk(ds1, -ds0*c1 - s0*dc1, dc0*c1 + c0*dc1) };
#pragma unroll
        for (int j = 1; j < 3; ++j) {
            SV HD;
            HD.w = mul(R_GF, HDwj[j-1]) + cross(w_GP, o.H[j].w);
            HD.v = mul(R_GF, cross(HDwj[j-1], r) + cross(k.Hw[j], wxr)) + cross(w_GP, o.H[j].v);
            VD = VD + u[j]*HD;
        }
    } else {
        SBK_ROTCOL_D(0, 0) SBK_ROTCOL_D(1, 1) SBK_ROTCOL_D(2, 2)
        if constexpr (isFreeLike(JT)) { SBK_TRCOL_D(3) SBK_TRCOL_D(4) SBK_TRCOL_D(5) }
    }
    #undef SBK_ROTCOL_D
    #undef SBK_TRCOL_D

    // ---- joint-independent velocity kinematics (RigidBodyNode.cpp:97-174) ----------------------
    o.V = phiT(o.l, V_GP) + V_PB;
    const V3 w = o.V.w;
    o.gyro.w = bc.mass*cross(w, mul(o.G, w)); o.gyro.v = bc.mass*cross(w, cross(w, o.c));
    o.acor.w = VD.w; o.acor.v = VD.v + cross(w_GP, o.V.v - v_GP);

    // ---- qdot = N(q) u  ------------------------------------------------------------------------
    if constexpr (isEulerKind(JT)) {              // qdot = N_P(q) w_FM (Rotation.h:395-406); unused slot = 0 (_Free.h:238-249)
        const V3 qd = eulerNTimes(k.sc, mk(u[0], u[1], u[2]));
        qdot[0] = qd.x; qdot[1] = qd.y; qdot[2] = qd.z;
        if constexpr (JT == JT_FREE_EULER) { qdot[3] = u[3]; qdot[4] = u[4]; qdot[5] = u[5]; qdot[6] = 0; }
        else qdot[3] = 0;
    } else if constexpr (JT == JT_BALL || JT == JT_FREE) {
        quatNTimes(q, mk(u[0], u[1], u[2]), qdot);
        if constexpr (JT == JT_FREE) {
#pragma unroll
            for (int i = 0; i < 3; ++i) qdot[4+i] = u[3+i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < d; ++i) qdot[i] = u[i];
    }
}

template <int JT>
SBK_HD void kinCore(const BodyConst& bc, const double* q, const double* u,
                    const M3& R_GP, const V3 p_GP, const SV V_GP,
                    KinOut<JointDims<JT>::nu>& o, double* qdot, double& qerr) {
    KinLocal<JointDims<JT>::nu> k;
    kinLocal<JT>(bc, q, k);
    kinGlobal<JT>(bc, k, q, u, R_GP, p_GP, V_GP, o, qdot);
    qerr = k.qerr;
}

// The recurrences X_GB = X_GP * X_PB and V_GB = ~Phi V_GP + H u run base->tip; the integrator
// path also needs them tip->base (the articulated-inertia sweep), and instead of storing every
// body's kinematics in HBM for that sweep it INVERTS the recurrence: given the body's own
// X_GB, V_GB it recovers the parent's,
//     R_GP = R_GB * ~R_PB,  p_GP = p_GB - R_GP p_PB,  V_GP = Phi^-T (V_GB - H u),
// and then recomputes the body's kinematic quantities from (X_GP, V_GP) with the same formulas
// as the outward sweeps.  Exact in exact arithmetic; in FP64 the recovered parent transform
// differs from the stored one by a few ulp per body (orthogonal factors, no amplification).
template <int JT>
SBK_HD void kinReverse(const BodyConst& bc, const KinLocal<JointDims<JT>::nu>& k, const double* u,
                       const M3& R_GB, const V3 p_GB, const SV V_GB, M3& R_GP, V3& p_GP, SV& V_GP) {
    constexpr int d = JointDims<JT>::nu;
    R_GP = mulABt(R_GB, k.R_PB);
    const V3 l = mul(R_GP, k.p_PB);
    p_GP = p_GB - l;
    const bool noRPF = (bc.flags & BF_NO_R_PF) != 0;
    const M3 R_GF = noRPF ? R_GP : mul(R_GP, loadR(bc.X_PF));
    SV H[dim1(d)];
    jointH<JT>(R_GF, k, H);
    SV V_PB = zeroSV();
    if constexpr (d > 0) V_PB = u[0]*H[0];
#pragma unroll
    for (int j = 1; j < d; ++j) V_PB = V_PB + u[j]*H[j];
    V_GP.w = V_GB.w - V_PB.w;
    V_GP.v = (V_GB.v - V_PB.v) - cross(V_GP.w, l);
}

template <int d> struct AbiOut { SV G[dim1(d)]; double DI[dim1(d*d)]; ABI PP; SV zb; bool ok; };

// Sweep C for one body, given P = Mk + sum of shifted children P+ (RigidBodyNodeSpec.cpp:249-325).
template <int d>
SBK_HD void abiCore(const ABI& P, const SV* H, const SV acor, const SV gyro, AbiOut<d>& o) {
    if constexpr (d == 0) {        // Weld: P+ = P (RigidBodyNode_Weld.cpp:437-450)
        o.PP = P; o.ok = true; o.zb = mul(P, acor) + gyro; return;
    } else {
    SV PH[dim1(d)];
#pragma unroll
    for (int j = 0; j < d; ++j) PH[j] = mul(P, H[j]);
    double D[dim1(d*d)];
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) D[d*i+j] = dot(H[i].w, PH[j].w) + dot(H[i].v, PH[j].v);
    o.ok = Inv<d>::run(D, o.DI);
#pragma unroll
    for (int j = 0; j < d; ++j) {
        SV g = zeroSV();
#pragma unroll
        for (int k = 0; k < d; ++k) g = g + o.DI[d*k+j]*PH[k];
        o.G[j] = g;
    }
    // PPlus = P - G*~PH, symmetrised (RigidBodyNodeSpec.cpp:309-324)
    double mm[9], ms[9], in[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { mm[i] = 0; ms[i] = 0; in[i] = 0; }
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const double gw[3] = {o.G[k].w.x, o.G[k].w.y, o.G[k].w.z}, gv[3] = {o.G[k].v.x, o.G[k].v.y, o.G[k].v.z};
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
    o.PP.M.xx = P.M.xx - ms[0]; o.PP.M.yy = P.M.yy - ms[4]; o.PP.M.zz = P.M.zz - ms[8];
    o.PP.M.xy = P.M.xy - (ms[3]+ms[1])/2; o.PP.M.xz = P.M.xz - (ms[6]+ms[2])/2; o.PP.M.yz = P.M.yz - (ms[7]+ms[5])/2;
    o.PP.J.xx = P.J.xx - in[0]; o.PP.J.yy = P.J.yy - in[4]; o.PP.J.zz = P.J.zz - in[8];
    o.PP.J.xy = P.J.xy - (in[3]+in[1])/2; o.PP.J.xz = P.J.xz - (in[6]+in[2])/2; o.PP.J.yz = P.J.yz - (in[7]+in[5])/2;
#pragma unroll
    for (int i = 0; i < 9; ++i) o.PP.F.a[i] = P.F.a[i] - mm[i];
    // realizeArticulatedBodyVelocityCache (RigidBodyNode.cpp:201-213): P*a + b
    o.zb = mul(P, acor) + gyro;
    }
}

// Sweep D for one body: z already holds (P a + b - F) + sum Phi*z+ of the children.
template <int d>
SBK_HD void zCore(const SV* H, const SV* G, const SV z, const double* f, double* eps, SV& zPlus) {
#pragma unroll
    for (int j = 0; j < d; ++j) eps[j] = f[j] - (dot(H[j].w, z.w) + dot(H[j].v, z.v));
    SV Ge = zeroSV();
#pragma unroll
    for (int j = 0; j < d; ++j) Ge = Ge + eps[j]*G[j];
    zPlus = z + Ge;
}

// Sweep E for one body.
template <int d, bool WITH_COR>
SBK_HD void accCore(const SV* H, const SV* G, const double* DI, const double* eps, const V3 l, const SV A_GP,
                    const SV acor, double* udot, SV& A) {
    const SV APlus = phiT(l, A_GP);
    SV Hu = zeroSV();
#pragma unroll
    for (int i = 0; i < d; ++i) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < d; ++j) s += DI[d*i+j]*eps[j];
        udot[i] = s - (dot(G[i].w, APlus.w) + dot(G[i].v, APlus.v));
        Hu = Hu + udot[i]*H[i];
    }
    A = APlus + Hu;
    if constexpr (WITH_COR) A = A + acor;
}

// calcQDotDot (RigidBodyNodeSpec_Ball.h:355-390, _Free.h:418-455)
template <int JT>
SBK_HD void qddCore(const double* q, const double* u, const double* udot, double* qdd) {
    constexpr int d = JointDims<JT>::nu;
    if constexpr (isEulerKind(JT)) {              // qdotdot = N b + NDot w (Rotation.h:1041-1060)
        double sc[5]; double s2unused, c2unused;
        sincos(q[0], &sc[1], &sc[0]); sincos(q[1], &sc[3], &sc[2]); sc[4] = 1/sc[2]; (void)s2unused; (void)c2unused;
        const V3 qd = eulerNTimes(sc, mk(u[0], u[1], u[2]));
        const V3 Nb = eulerNTimes(sc, mk(udot[0], udot[1], udot[2]));
        const double q1oc1 = qd.y*sc[4];
        qdd[0] = Nb.x + (qd.x*sc[3] - qd.z)*q1oc1; qdd[1] = Nb.y + qd.x*qd.z*sc[2]; qdd[2] = Nb.z + (qd.z*sc[3] - qd.x)*q1oc1;
        if constexpr (JT == JT_FREE_EULER) { qdd[3] = udot[3]; qdd[4] = udot[4]; qdd[5] = udot[5]; qdd[6] = 0; }
        else qdd[3] = 0;
    } else if constexpr (JT == JT_BALL || JT == JT_FREE) {
        const V3 w = mk(u[0], u[1], u[2]);
        double Nb[4]; quatNTimes(q, mk(udot[0], udot[1], udot[2]), Nb);
        const double k = -0.25*dot(w, w);
#pragma unroll
        for (int i = 0; i < 4; ++i) qdd[i] = Nb[i] + k*q[i];
        if constexpr (JT == JT_FREE) {
#pragma unroll
            for (int i = 0; i < 3; ++i) qdd[4+i] = udot[3+i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < d; ++i) qdd[i] = udot[i];
    }
}

// Force::Gravity on one body (Force_Gravity.cpp:538-556): F = (p_CB_G x m g, m g)
SBK_HD SV gravityForce(const double mass, const V3 c_G, const double gx, const double gy, const double gz) {
    const V3 Fc = mass*mk(gx, gy, gz);
    SV F; F.w = cross(c_G, Fc); F.v = Fc; return F;
}
// MobilityLinearSpring / Damper / MobilityConstantForce of one body in force-index order (Force.cpp:339-351,434-443,
// Force_MobilityConstantForce.cpp);
// q, u are the body's own coordinates.
template <int d>
SBK_HD void mobilityForces(const BodyConst& bc, const ForceConst* forces, const double* q, const double* u, double* f) {
#pragma unroll
    for (int j = 0; j < d; ++j) f[j] = 0;
    for (int k = 0; k < bc.nforce; ++k) {
        const ForceConst fc = forces[bc.forceStart + k];
#pragma unroll
        for (int j = 0; j < d; ++j) if (j == fc.coord) {
            const double frc = (fc.kind == FK_SPRING) ? -fc.a*(q[j] - fc.b) : (fc.kind == FK_CONSTANT) ? fc.a : -fc.a*u[j];
            f[j] += frc;
        }
    }
}

//==============================================================================================
//                     BODY WRAPPERS, FULL records (API realize path)
//==============================================================================================
// Every field of the per-body record is stored: getters and the operator forms need them.
// CB = true: the records are CTA-blocked (plans 1 and 4 on the device) and the row stride is the compile-time 128.

template <int JT, bool CB = false>
SBK_BODY void kinBody(const Ctx& c, const BodyConst& bc, const int bodyIndex, const int inst, double* qdotDst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    const CacheRefT<CB> me = cacheOf<CB>(c, inst, bc.cacheBase);
    double q[dim1(NQ)], u[dim1(d)], qdot[dim1(NQ)], qerr;
#pragma unroll
    for (int i = 0; i < NQ; ++i) q[i] = ldS<false>(c, inst, c.q, bc.q0 + i);
#pragma unroll
    for (int i = 0; i < d; ++i)  u[i] = ldS<false>(c, inst, c.u, bc.u0 + i);

    const CacheRefT<CB> pa = cacheOf<CB>(c, inst, bc.parentCacheBase);
    const M3 R_GP = pa.ldM3(F_XGB); const V3 p_GP = pa.ld3(F_XGB + 9); const SV V_GP = pa.ldSV(F_VGB);

    KinOut<d> o;
    kinCore<JT>(bc, q, u, R_GP, p_GP, V_GP, o, qdot, qerr);

    me.stM3(F_XGB, o.R); me.st3(F_XGB + 9, o.p); me.stSV(F_VGB, o.V);
    me.st3(F_L, o.l); me.st3(F_MK, o.c); me.stS3(F_MK + 3, o.G);
    me.stSV(F_ACOR, o.acor); me.stSV(F_GYRO, o.gyro);
#pragma unroll
    for (int j = 0; j < d; ++j) me.stSV(F_H + 6*j, o.H[j]);
    if (qdotDst) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) stS<false>(c, inst, qdotDst, bc.q0 + i, qdot[i]);
    }
    if constexpr (JT == JT_BALL || JT == JT_FREE) { if (c.qerr) stS<false>(c, inst, c.qerr, bc.quat, qerr); }
}

// Force::TwoPointLinearSpringImpl::calcForce / TwoPointLinearDamperImpl::calcForce (Force.cpp:103-140,179-221) for one instance from the
// realized position / velocity records: f2 (zeroed here) receives +(s1_G x f, f) on body1 and -(s2_G x f, f) on body2, in element order.
template <bool CB = false>
SBK_HD void twoPointPass(const Ctx& c, const int inst) {
    if (c.ntp == 0) return;
    for (int i = 0; i < 6*c.nb; ++i) stS<false>(c, inst, c.f2, i, 0.0);
    for (int e = 0; e < c.ntp; ++e) {
        const TwoPointConst& t = c.tp[e];
        const CacheRefT<CB> r1 = cacheOf<CB>(c, inst, t.cacheBase1), r2 = cacheOf<CB>(c, inst, t.cacheBase2);
        const V3 s1_G = mul(r1.ldM3(F_XGB), mk(t.s1[0], t.s1[1], t.s1[2])), s2_G = mul(r2.ldM3(F_XGB), mk(t.s2[0], t.s2[1], t.s2[2]));
        const V3 p1_G = r1.ld3(F_XGB + 9) + s1_G, p2_G = r2.ld3(F_XGB + 9) + s2_G;
        const V3 r_G = p2_G - p1_G;
        const double d = sqrt(dot(r_G, r_G));
        V3 f1;
        if (t.kind == TP_SPRING) { const double frc = t.a*(d - t.b); f1 = (frc/d)*r_G; }
        else {      // findStationVelocityInGround: v + w x s_G; UnitVec3(r) = r / |r|
            const SV V1 = r1.ldSV(F_VGB), V2 = r2.ldSV(F_VGB);
            const V3 vRel = (V2.v + cross(V2.w, s2_G)) - (V1.v + cross(V1.w, s1_G));
            const V3 dir = (1.0/d)*r_G;
            f1 = (t.a*dot(vRel, dir))*dir;
        }
        const V3 m1 = cross(s1_G, f1), m2 = cross(s2_G, f1);
        const double add1[6] = {m1.x, m1.y, m1.z, f1.x, f1.y, f1.z}, add2[6] = {m2.x, m2.y, m2.z, f1.x, f1.y, f1.z};
        for (int i = 0; i < 6; ++i) {
            stS<false>(c, inst, c.f2, 6*t.body1 + i, ldS<false>(c, inst, c.f2, 6*t.body1 + i) + add1[i]);
            stS<false>(c, inst, c.f2, 6*t.body2 + i, ldS<false>(c, inst, c.f2, 6*t.body2 + i) - add2[i]);
        }
    }
}

// Inward body step.  MODE bits:
//   IN_ABI    articulated-body inertia: P, D, DI, G, PPlus (+ ZB = P*a + b)
//   IN_Z      residual pass: z, eps, zPlus
//   IN_BIAS   z starts from ZB - F (forward dynamics); otherwise from 0 (M^-1)
//   IN_FORCES applied forces come from the lowered force elements (gravity/spring/damper)
//             rather than from c.fmobIn / c.FbodyIn
enum { IN_ABI = 1, IN_Z = 2, IN_BIAS = 4, IN_FORCES = 8 };

SBK_HD void setSingular(const Ctx& c, const int inst) {
    if (!c.status) return;
#if defined(__CUDA_ARCH__)
    atomicOr(c.status + inst, 2);
#else
    c.status[inst] |= 2;
#endif
}

template <int JT, int MODE, bool CB = false>
SBK_BODY void inwardBody(const Ctx& c, const BodyConst& bc, const int bodyIndex, const int inst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    const CacheRefT<CB> me = cacheOf<CB>(c, inst, bc.cacheBase);
    SV H[dim1(d)];
#pragma unroll
    for (int j = 0; j < d; ++j) H[j] = me.ldSV(F_H + 6*j);
    const V3 c_G = me.ld3(F_MK);

    AbiOut<d> ao; ao.zb = zeroSV();
    if constexpr ((MODE & IN_ABI) != 0) {
        const S3 G_G = me.ldS3(F_MK + 3); const SV acor = me.ldSV(F_ACOR), gyro = me.ldSV(F_GYRO);
        ABI P = abiFromRigid(bc.mass, c_G, G_G);
        for (int k = 0; k < bc.nchild; ++k) {
            const CacheRefT<CB> ch = cacheOf<CB>(c, inst, c.bodies[c.children[bc.childStart + k]].cacheBase);
            addInto(P, shiftABI(ch.ldABI(F_PPLUS), ch.ld3(F_L)));
        }
        abiCore<d>(P, H, acor, gyro, ao);
        if (!ao.ok) setSingular(c, inst);
#pragma unroll
        for (int j = 0; j < d; ++j) me.stSV(fG(d) + 6*j, ao.G[j]);
#pragma unroll
        for (int i = 0; i < d*d; ++i) me.st(fDI(d) + i, ao.DI[i]);
        me.stABI(F_PPLUS, ao.PP);
        me.stSV(F_ZB, ao.zb);
    } else {
#pragma unroll
        for (int j = 0; j < d; ++j) ao.G[j] = me.ldSV(fG(d) + 6*j);
        if constexpr ((MODE & IN_BIAS) != 0) ao.zb = me.ldSV(F_ZB);
    }

    if constexpr ((MODE & IN_Z) != 0) {
        // ---- applied forces -------------------------------------------------------------------
        SV F = zeroSV(); double f[dim1(d)];
        if constexpr ((MODE & IN_FORCES) != 0) {
            double qF[dim1(NQ)], uF[dim1(d)];
            if (bc.nforce > 0) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) qF[i] = ldS<false>(c, inst, c.q, bc.q0 + i);
#pragma unroll
                for (int i = 0; i < d; ++i)  uF[i] = ldS<false>(c, inst, c.u, bc.u0 + i);
            }
            F = gravityForce(bc.mass, c_G, c.gx, c.gy, c.gz);
            if (c.ntp) {
                F.w = F.w + mk(ldS<false>(c, inst, c.f2, 6*bodyIndex+0), ldS<false>(c, inst, c.f2, 6*bodyIndex+1), ldS<false>(c, inst, c.f2, 6*bodyIndex+2));
                F.v = F.v + mk(ldS<false>(c, inst, c.f2, 6*bodyIndex+3), ldS<false>(c, inst, c.f2, 6*bodyIndex+4), ldS<false>(c, inst, c.f2, 6*bodyIndex+5));
            }
            mobilityForces<d>(bc, c.forces, qF, uF, f);
            if (c.fmobOut) {
#pragma unroll
                for (int j = 0; j < d; ++j) stS<false>(c, inst, c.fmobOut, bc.u0 + j, f[j]);
            }
            if (c.FbodyOut) {
                stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+0, F.w.x); stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+1, F.w.y); stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+2, F.w.z);
                stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+3, F.v.x); stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+4, F.v.y); stS<false>(c, inst, c.FbodyOut, 6*bodyIndex+5, F.v.z);
            }
        } else {
#pragma unroll
            for (int j = 0; j < d; ++j) f[j] = c.fmobIn ? ldS<false>(c, inst, c.fmobIn, bc.u0 + j) : 0.0;
            if (c.FbodyIn) {
                F.w = mk(ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+0), ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+1), ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+2));
                F.v = mk(ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+3), ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+4), ldS<false>(c, inst, c.FbodyIn, 6*bodyIndex+5));
            }
        }
        // ---- calcUDotPass1Inward (RigidBodyNodeSpec.cpp:355-400) / M^-1 pass 1 (:483-515) ------
        SV z;
        if constexpr ((MODE & IN_BIAS) != 0) z = ao.zb - F; else z = zeroSV();
        for (int k = 0; k < bc.nchild; ++k) {
            const CacheRefT<CB> ch = cacheOf<CB>(c, inst, c.bodies[c.children[bc.childStart + k]].cacheBase);
            z = z + phi(ch.ld3(F_L), ch.ldSV(F_ZPLUS));
        }
        double eps[dim1(d)]; SV zPlus;
        zCore<d>(H, ao.G, z, f, eps, zPlus);
#pragma unroll
        for (int j = 0; j < d; ++j) me.st(fEPS(d) + j, eps[j]);
        me.stSV(F_ZPLUS, zPlus);
    }
}

// Sweep E for one body (base->tip): udot, A_GB, qdotdot.
template <int JT, bool WITH_COR, bool CB = false>
SBK_BODY void outwardBody(const Ctx& c, const BodyConst& bc, const int bodyIndex, const int inst,
                         double* udotDst, double* qdotdotDst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    const CacheRefT<CB> me = cacheOf<CB>(c, inst, bc.cacheBase);
    const SV A_GP = cacheOf<CB>(c, inst, bc.parentCacheBase).ldSV(F_AGB);
    SV H[dim1(d)], G[dim1(d)]; double DI[dim1(d*d)], eps[dim1(d)], udot[dim1(d)];
#pragma unroll
    for (int j = 0; j < d; ++j) { H[j] = me.ldSV(F_H + 6*j); G[j] = me.ldSV(fG(d) + 6*j); eps[j] = me.ld(fEPS(d) + j); }
#pragma unroll
    for (int i = 0; i < d*d; ++i) DI[i] = me.ld(fDI(d) + i);
    SV acor = zeroSV();
    if constexpr (WITH_COR) acor = me.ldSV(F_ACOR);
    const V3 lMe = me.ld3(F_L);
    SV A;
    accCore<d, WITH_COR>(H, G, DI, eps, lMe, A_GP, acor, udot, A);
    me.stSV(F_AGB, A);
    if (udotDst) {
#pragma unroll
        for (int i = 0; i < d; ++i) stS<false>(c, inst, udotDst, bc.u0 + i, udot[i]);
    }
    if (qdotdotDst) {
        double q[dim1(NQ)], u[dim1(d)], qdd[dim1(NQ)];
        if constexpr (JT == JT_BALL || JT == JT_FREE || isEulerKind(JT)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = ldS<false>(c, inst, c.q, bc.q0 + i);
#pragma unroll
            for (int i = 0; i < 3; ++i) u[i] = ldS<false>(c, inst, c.u, bc.u0 + i);
        }
        qddCore<JT>(q, u, udot, qdd);
#pragma unroll
        for (int i = 0; i < NQ; ++i) stS<false>(c, inst, qdotdotDst, bc.q0 + i, qdd[i]);
    }
}

//==============================================================================================
//        BODY WRAPPERS, integrator path (LEAN): reversible kinematics, 7*dof doubles per body
//==============================================================================================
// One derivative evaluation = three sweeps over the bodies of one instance:
//   1 outward  position/velocity recurrences only; X_GB, V_GB ride in the carry and are stored
//              just where a later sweep cannot get them from the carry (tips, branch points)
//   2 inward   per body: recover the parent's X, V from the body's own (kinReverse), recompute the
//              body's kinematics, then articulated inertia + residual; stores G (6 dof) and
//              nu = DI*eps (dof)
//   3 outward  recompute the kinematics (bit-identical to sweep 1), udot = nu - ~G A+, A_GB
// HBM traffic per body and evaluation: 7*dof doubles written + read (Pin: 112 B) instead of the
// 728 B of a design that stages every body's kinematics between the sweeps; the price is ~2x the
// algorithmic flop count, paid on an FP64 pipe that the staged design left 85% idle.
// LEAN record rows (inside the body's FULL record region, which is larger):

enum { LF_XGB = 0, LF_VGB = 12, LF_L = 18, LF_PPLUS = 21, LF_ZPLUS = 42, LF_AGB = 48, LF_G = 54 };
SBK_HD constexpr int lfNU(int d) { return LF_G + 6*d; }

// Coordinates of the NEXT body of a sweep, requested one body step ahead: every body step starts
// with sin/cos of q, so a load issued at its top stalls the step for a full trip to memory.  The
// request is an asynchronous global->shared copy (cp.async, SASS LDGSTS) into two alternating
// 4-row slots of the carry column -- registers would not do: the compiler spills a value that is
// live across the joint switch, and the spill store waits for the load.  Two slots each for q and
// u cover Pin / Slider / Universal; Ball / Free bodies load their own.
template <int JMASK>
SBK_HD void preloadCoords(const Ctx& c, const int inst, const BodyConst& nx, double* slot) {
    constexpr bool BLK = SBK_DEV_BLK;
    // a model made of 1-dof mobilizers only needs one q and one u row per body (every cp.async costs: measured)
    constexpr bool ONE = (JMASK & ~(JM_PIN | JM_SLIDER | JM_WELD)) == 0;
    // rows of [q; u] (the u rows follow the q rows), clamped: a Weld owns no slots and the slot after the
    // last one does not exist
    const int last = c.nq + c.nu - 1;
    // a quaternion mobilizer gets its four quaternion rows (the body step starts with |q|; its u is needed later and
    // loaded directly); every other kind its first two q and two u rows
    const bool quat = (JMASK & (JM_BALL | JM_FREE)) != 0 && (nx.joint == JT_BALL || nx.joint == JT_FREE);
    const int r0 = nx.q0 < last ? nx.q0 : last, r1 = nx.q0 + 1 < last ? nx.q0 + 1 : last;
    const int r2 = quat ? nx.q0 + 2 : (c.nq + nx.u0 < last ? c.nq + nx.u0 : last), r3 = quat ? nx.q0 + 3 : (c.nq + nx.u0 + 1 < last ? c.nq + nx.u0 + 1 : last);
    const double* src[4] = { c.q + stateIndex<BLK>(c, inst, r0), c.q + stateIndex<BLK>(c, inst, r1),
                             c.q + stateIndex<BLK>(c, inst, r2), c.q + stateIndex<BLK>(c, inst, r3) };
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (!ONE || k == 0 || k == 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(slot + k*SBK_CARRY_STRIDE)), "l"(src[k]) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    for (int k = 0; k < 4; ++k) slot[k*SBK_CARRY_STRIDE] = *src[k];
#endif
}
// G and nu = DI*eps of the next body of the acceleration sweep (7*dof contiguous record rows, dof <= 2),
// requested together with its coordinates.  Joins the commit group of the following preloadCoords.
SBK_HD void preloadGNu(const Ctx& c, const int inst, const BodyConst& nx, double* slot) {
    constexpr bool BLK = SBK_DEV_BLK;
    const int d = (nx.joint == JT_UNIVERSAL) ? 2 : 1;
    if (nx.joint > JT_UNIVERSAL) return;
    const CacheRefT<BLK> rec = cacheOf<BLK>(c, inst, nx.cacheBase);
    for (int k = 0; k < 7*d; ++k) {
        const double* src = rec.p + (BLK ? (long long)(LF_G + k)*BLK_LANES : (long long)(LF_G + k)*rec.stride);
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(slot + k*SBK_CARRY_STRIDE)), "l"(src) : "memory");
#else
        slot[k*SBK_CARRY_STRIDE] = *src;
#endif
    }
}
// Wait until every request but the most recent one has landed.
SBK_HD void preloadWait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
template <int JT, bool BLK> SBK_HD void takeCoords(const Ctx& c, const int inst, const BodyConst& bc, const double* slot, double* q, double* u) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    if constexpr (NQ <= 2) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[i] = slot[i*SBK_CARRY_STRIDE];
#pragma unroll
        for (int i = 0; i < d; ++i)  u[i] = slot[(2 + i)*SBK_CARRY_STRIDE];
    } else if constexpr (JT == JT_BALL || JT == JT_FREE) {     // the quaternion was preloaded, the rest is needed later
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = slot[i*SBK_CARRY_STRIDE];
#pragma unroll
        for (int i = 4; i < NQ; ++i) q[i] = ldS<BLK>(c, inst, c.q, bc.q0 + i);
#pragma unroll
        for (int i = 0; i < d; ++i)  u[i] = ldS<BLK>(c, inst, c.u, bc.u0 + i);
    } else {
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[i] = ldS<BLK>(c, inst, c.q, bc.q0 + i);
#pragma unroll
        for (int i = 0; i < d; ++i)  u[i] = ldS<BLK>(c, inst, c.u, bc.u0 + i);
    }
}

template <int JT>
SBK_BODY void leanKinBody(const Ctx& c, const BodyConst& bc, const int inst, double* cy, const double* pre, double* qdotDst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    double q[dim1(NQ)], u[dim1(d)], qdot[dim1(NQ)], qerr;
    takeCoords<JT, BLK>(c, inst, bc, pre, q, u);
    M3 R_GP; V3 p_GP; SV V_GP;
    if (bc.flags & BF_PARENT_PREV) cyLoadOut(cy, R_GP, p_GP, V_GP);
    else { const CacheRefT<BLK> pa = cacheOf<BLK>(c, inst, bc.parentCacheBase); R_GP = pa.ldM3(LF_XGB); p_GP = pa.ld3(LF_XGB + 9); V_GP = pa.ldSV(LF_VGB); }
    KinOut<d> o;
    kinCore<JT>(bc, q, u, R_GP, p_GP, V_GP, o, qdot, qerr);     // unused outputs are dead code here
    if (bc.flags & (BF_STORE_LINK | BF_TIP)) {
        const CacheRefT<BLK> me = cacheOf<BLK>(c, inst, bc.cacheBase);
        me.stM3(LF_XGB, o.R); me.st3(LF_XGB + 9, o.p); me.stSV(LF_VGB, o.V);
    }
    cyStoreOut(cy, o.R, o.p, o.V);
    if (qdotDst) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) stS<BLK>(c, inst, qdotDst, bc.q0 + i, qdot[i]);
    }
}

template <int JT>
SBK_BODY void leanInwardBody(const Ctx& c, const Tables& T, const BodyConst& bc, const int bodyIndex, const int inst, double* cy, const double* pre) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    const CacheRefT<BLK> me = cacheOf<BLK>(c, inst, bc.cacheBase);
    double q[dim1(NQ)], u[dim1(d)], qdot[dim1(NQ)];
    takeCoords<JT, BLK>(c, inst, bc, pre, q, u);
    const bool haveCarryChild = !(bc.flags & BF_TIP);      // body index + 1 is a child and left its links in the carry
    M3 R_GB; V3 p_GB; SV V_GB;
    if (haveCarryChild) cyLoadOut(cy + CY_SELF*SBK_CARRY_STRIDE, R_GB, p_GB, V_GB);
    else { R_GB = me.ldM3(LF_XGB); p_GB = me.ld3(LF_XGB + 9); V_GB = me.ldSV(LF_VGB); }

    KinLocal<d> kl; kinLocal<JT>(bc, q, kl);
    M3 R_GP; V3 p_GP; SV V_GP;
    kinReverse<JT>(bc, kl, u, R_GB, p_GB, V_GB, R_GP, p_GP, V_GP);
    KinOut<d> o;
    kinGlobal<JT>(bc, kl, q, u, R_GP, p_GP, V_GP, o, qdot);

    // ---- articulated inertia (RigidBodyNodeSpec.cpp:249-325) ------------------------------------
    ABI cPP; SV czP = zeroSV(); V3 cl = zero3();
    if (haveCarryChild) cyLoadIn(cy, cPP, czP, cl);
    ABI P = abiFromRigid(bc.mass, o.c, o.G);
    for (int k = 0; k < bc.nchild; ++k) {
        if (k == 0 && haveCarryChild) { addInto(P, shiftABI(cPP, cl)); continue; }
        const CacheRefT<BLK> ch = cacheOf<BLK>(c, inst, T.bodies[T.children[bc.childStart + k]].cacheBase);
        addInto(P, shiftABI(ch.ldABI(LF_PPLUS), ch.ld3(LF_L)));
    }
    AbiOut<d> ao;
    abiCore<d>(P, o.H, o.acor, o.gyro, ao);
    if (!ao.ok) setSingular(c, inst);

    // ---- forces and residual (RigidBodyNodeSpec.cpp:355-400) -------------------------------------
    double f[dim1(d)];
    const SV F = gravityForce(bc.mass, o.c, c.gx, c.gy, c.gz);
    mobilityForces<d>(bc, T.forces, q, u, f);
    SV z = ao.zb - F;
    for (int k = 0; k < bc.nchild; ++k) {
        if (k == 0 && haveCarryChild) { z = z + phi(cl, czP); continue; }
        const CacheRefT<BLK> ch = cacheOf<BLK>(c, inst, T.bodies[T.children[bc.childStart + k]].cacheBase);
        z = z + phi(ch.ld3(LF_L), ch.ldSV(LF_ZPLUS));
    }
    double eps[dim1(d)]; SV zPlus;
    zCore<d>(o.H, ao.G, z, f, eps, zPlus);
#pragma unroll
    for (int j = 0; j < d; ++j) me.stSV(LF_G + 6*j, ao.G[j]);
#pragma unroll
    for (int i = 0; i < d; ++i) {                      // nu = DI*eps, the first term of udot (RigidBodyNodeSpec.cpp:432)
        double sum = 0;
#pragma unroll
        for (int j = 0; j < d; ++j) sum += ao.DI[d*i+j]*eps[j];
        me.st(lfNU(d) + i, sum);
    }
    if (bc.flags & BF_PARENT_PREV) {
        cyStoreIn(cy, ao.PP, zPlus, o.l);
        cyStoreOut(cy + CY_SELF*SBK_CARRY_STRIDE, R_GP, p_GP, V_GP);
    } else {
        me.stABI(LF_PPLUS, ao.PP); me.stSV(LF_ZPLUS, zPlus); me.st3(LF_L, o.l);
    }
}

template <int JT>
SBK_BODY void leanOutwardBody(const Ctx& c, const BodyConst& bc, const int inst, double* cy, const double* pre, const double* gnu, double* udotDst, double* qdotdotDst) {
    constexpr int NQ = JointDims<JT>::nq, d = JointDims<JT>::nu;
    constexpr bool BLK = SBK_DEV_BLK;
    const CacheRefT<BLK> me = cacheOf<BLK>(c, inst, bc.cacheBase);
    double q[dim1(NQ)], u[dim1(d)], qdot[dim1(NQ)], qerr, nu[dim1(d)], udot[dim1(d)];
    SV G[dim1(d)];
    if ((JT == JT_PIN || JT == JT_SLIDER || JT == JT_UNIVERSAL) && gnu) {   // preloaded into the carry column (preloadGNu)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            G[j].w = mk(gnu[(6*j+0)*SBK_CARRY_STRIDE], gnu[(6*j+1)*SBK_CARRY_STRIDE], gnu[(6*j+2)*SBK_CARRY_STRIDE]);
            G[j].v = mk(gnu[(6*j+3)*SBK_CARRY_STRIDE], gnu[(6*j+4)*SBK_CARRY_STRIDE], gnu[(6*j+5)*SBK_CARRY_STRIDE]);
            nu[j] = gnu[(6*d+j)*SBK_CARRY_STRIDE];
        }
    } else {
#pragma unroll
        for (int j = 0; j < d; ++j) { G[j] = me.ldSV(LF_G + 6*j); nu[j] = me.ld(lfNU(d) + j); }
    }
    takeCoords<JT, BLK>(c, inst, bc, pre, q, u);
    M3 R_GP; V3 p_GP; SV V_GP, A_GP;
    if (bc.flags & BF_PARENT_PREV) { cyLoadOut(cy, R_GP, p_GP, V_GP); A_GP = cyLoadA(cy + CY_A*SBK_CARRY_STRIDE); }
    else {
        const CacheRefT<BLK> pa = cacheOf<BLK>(c, inst, bc.parentCacheBase);
        R_GP = pa.ldM3(LF_XGB); p_GP = pa.ld3(LF_XGB + 9); V_GP = pa.ldSV(LF_VGB); A_GP = pa.ldSV(LF_AGB);
    }
    KinOut<d> o;
    kinCore<JT>(bc, q, u, R_GP, p_GP, V_GP, o, qdot, qerr);     // mass properties are dead code here
    // calcUDotPass2Outward (RigidBodyNodeSpec.cpp:408-446) with nu = DI*eps from the inward sweep
    const SV APlus = phiT(o.l, A_GP);
    SV Hu = zeroSV();
#pragma unroll
    for (int i = 0; i < d; ++i) {
        udot[i] = nu[i] - (dot(G[i].w, APlus.w) + dot(G[i].v, APlus.v));
        Hu = Hu + udot[i]*o.H[i];
    }
    const SV A = (APlus + Hu) + o.acor;
    if (bc.flags & BF_STORE_LINK) me.stSV(LF_AGB, A);            // X_GB, V_GB are still there from sweep 1
    cyStoreOut(cy, o.R, o.p, o.V); cyStoreA(cy + CY_A*SBK_CARRY_STRIDE, A);
    if (udotDst) {
#pragma unroll
        for (int i = 0; i < d; ++i) stS<BLK>(c, inst, udotDst, bc.u0 + i, udot[i]);
    }
    if (qdotdotDst) {
        double qdd[dim1(NQ)];
        qddCore<JT>(q, u, udot, qdd);
#pragma unroll
        for (int i = 0; i < NQ; ++i) stS<BLK>(c, inst, qdotdotDst, bc.q0 + i, qdd[i]);
    }
}

// multiplyByM / inverse dynamics passes (RigidBodyNodeSpec.cpp:566-695).
//   WITH_VEL: residual form (adds coriolis a, gyroscopic b, applied forces).
// The outward pass stores A in AGB; the inward pass stores F in ZPLUS.
template <int JT, bool WITH_VEL, bool CB = false>
SBK_BODY void idOutBody(const Ctx& c, const BodyConst& bc, const int inst) {
    constexpr int d = JointDims<JT>::nu;
    constexpr bool BLK = false;
    const CacheRefT<CB> me = cacheOf<CB>(c, inst, bc.cacheBase), pa = cacheOf<CB>(c, inst, bc.parentCacheBase);
    SV A = phiT(me.ld3(F_L), pa.ldSV(F_AGB));
    SV Hu = zeroSV();
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double v = c.vecIn ? ldS<BLK>(c, inst, c.vecIn, bc.u0 + j) : 0.0;
        Hu = Hu + v*me.ldSV(F_H + 6*j);
    }
    A = A + Hu;
    if constexpr (WITH_VEL) A = A + me.ldSV(F_ACOR);
    me.stSV(F_AGB, A);
}
template <int JT, bool WITH_VEL, bool CB = false>
SBK_BODY void idInBody(const Ctx& c, const BodyConst& bc, const int bodyIndex, const int inst) {
    constexpr int d = JointDims<JT>::nu;
    constexpr bool BLK = false;
    const CacheRefT<CB> me = cacheOf<CB>(c, inst, bc.cacheBase);
    const V3 c_G = me.ld3(F_MK); const S3 G_G = me.ldS3(F_MK + 3);
    SV F = mulSpatialInertia(bc.mass, c_G, G_G, me.ldSV(F_AGB));
    if constexpr (WITH_VEL) {
        F = F + me.ldSV(F_GYRO);
        if (c.FbodyIn) {
            SV Fa;
            Fa.w = mk(ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+0), ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+1), ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+2));
            Fa.v = mk(ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+3), ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+4), ldS<BLK>(c, inst, c.FbodyIn, 6*bodyIndex+5));
            F = F - Fa;
        }
    }
    for (int k = 0; k < bc.nchild; ++k) {
        const BodyConst& cb = c.bodies[c.children[bc.childStart + k]];
        const CacheRefT<CB> ch = cacheOf<CB>(c, inst, cb.cacheBase);
        F = F + phi(ch.ld3(F_L), ch.ldSV(F_ZPLUS));
    }
    me.stSV(F_ZPLUS, F);
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const SV Hj = me.ldSV(F_H + 6*j);
        double tau = dot(Hj.w, F.w) + dot(Hj.v, F.v);
        if constexpr (WITH_VEL) { if (c.fmobIn) tau -= ldS<BLK>(c, inst, c.fmobIn, bc.u0 + j); }
        stS<BLK>(c, inst, c.vecOut, bc.u0 + j, tau);
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
        case JT_WELD:      { constexpr int JT = JT_WELD;      CALL; } break;        \
        case JT_TRANSLATION: { constexpr int JT = JT_TRANSLATION; CALL; } break;    \
        case JT_CYLINDER:  { constexpr int JT = JT_CYLINDER;  CALL; } break;        \
        case JT_PLANAR:    { constexpr int JT = JT_PLANAR;    CALL; } break;        \
        case JT_GIMBAL:    { constexpr int JT = JT_GIMBAL;    CALL; } break;        \
        case JT_BALL_EULER: { constexpr int JT = JT_BALL_EULER; CALL; } break;      \
        case JT_FREE_EULER: { constexpr int JT = JT_FREE_EULER; CALL; } break;      \
        default: break;                                                             \
    }

template <bool CB = false> SBK_HD void kinDispatch(const Ctx& c, int b, int inst, double* qdotDst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (kinBody<JT, CB>(c, bc, b, inst, qdotDst)));
}
template <int MODE, bool CB = false> SBK_HD void inwardDispatch(const Ctx& c, int b, int inst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (inwardBody<JT, MODE, CB>(c, bc, b, inst)));
}
template <bool WITH_COR, bool CB = false> SBK_HD void outwardDispatch(const Ctx& c, int b, int inst, double* udotDst, double* qddDst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (outwardBody<JT, WITH_COR, CB>(c, bc, b, inst, udotDst, qddDst)));
}
// Same, restricted at compile time to the mobilizer kinds in JMASK (bit JT_x set = kind present): the
// integrator kernels are instantiated for a few masks so that a model made of Pin joints only does
// not carry (and register-allocate for) the Ball / Free code.
#define SBK_DISPATCH_JOINT_M(JMASK, jt, CALL)                                                                       \
    switch (jt) {                                                                                                   \
        case JT_PIN:       if constexpr (((JMASK) & JM_PIN) != 0)       { constexpr int JT = JT_PIN;       CALL; } break; \
        case JT_SLIDER:    if constexpr (((JMASK) & JM_SLIDER) != 0)    { constexpr int JT = JT_SLIDER;    CALL; } break; \
        case JT_UNIVERSAL: if constexpr (((JMASK) & JM_UNIVERSAL) != 0) { constexpr int JT = JT_UNIVERSAL; CALL; } break; \
        case JT_BALL:      if constexpr (((JMASK) & JM_BALL) != 0)      { constexpr int JT = JT_BALL;      CALL; } break; \
        case JT_FREE:      if constexpr (((JMASK) & JM_FREE) != 0)      { constexpr int JT = JT_FREE;      CALL; } break; \
        case JT_WELD:      if constexpr (((JMASK) & JM_WELD) != 0)      { constexpr int JT = JT_WELD;      CALL; } break; \
        case JT_TRANSLATION: if constexpr (((JMASK) & JM_TRANSLATION) != 0) { constexpr int JT = JT_TRANSLATION; CALL; } break; \
        case JT_CYLINDER:  if constexpr (((JMASK) & JM_CYLINDER) != 0)  { constexpr int JT = JT_CYLINDER;  CALL; } break; \
        case JT_PLANAR:    if constexpr (((JMASK) & JM_PLANAR) != 0)    { constexpr int JT = JT_PLANAR;    CALL; } break; \
        case JT_GIMBAL:    if constexpr (((JMASK) & JM_GIMBAL) != 0)    { constexpr int JT = JT_GIMBAL;    CALL; } break; \
        case JT_BALL_EULER: if constexpr (((JMASK) & JM_BALL_EULER) != 0) { constexpr int JT = JT_BALL_EULER; CALL; } break; \
        case JT_FREE_EULER: if constexpr (((JMASK) & JM_FREE_EULER) != 0) { constexpr int JT = JT_FREE_EULER; CALL; } break; \
        default: break;                                                                                             \
    }

template <bool WITH_VEL, bool CB = false> SBK_HD void idOutDispatch(const Ctx& c, int b, int inst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (idOutBody<JT, WITH_VEL, CB>(c, bc, inst)));
}
template <bool WITH_VEL, bool CB = false> SBK_HD void idInDispatch(const Ctx& c, int b, int inst) {
    const BodyConst& bc = c.bodies[b];
    SBK_DISPATCH_JOINT(bc.joint, (idInBody<JT, WITH_VEL, CB>(c, bc, b, inst)));
}

//==============================================================================================
// Per-instance drivers for the thread-per-instance plan: body index order is a valid
// base->tip order because a parent's MobilizedBodyIndex is always smaller than its child's.
//==============================================================================================
template <bool CB = false> SBK_HD void tpiKinematics(const Ctx& c, int inst, double* qdotDst) {
    for (int b = 1; b < c.nb; ++b) kinDispatch<CB>(c, b, inst, qdotDst);
}
template <int MODE, bool CB = false> SBK_HD void tpiInward(const Ctx& c, int inst) {
    for (int b = c.nb - 1; b >= 1; --b) inwardDispatch<MODE, CB>(c, b, inst);
}
template <bool WITH_COR, bool CB = false> SBK_HD void tpiOutward(const Ctx& c, int inst, double* udotDst, double* qddDst) {
    for (int b = 1; b < c.nb; ++b) outwardDispatch<WITH_COR, CB>(c, b, inst, udotDst, qddDst);
}
// One full derivative evaluation = System::realize(Acceleration) for the lowered system.
//   LEAN = false: FULL records (every cache entry a getter may ask for); cy unused
//   LEAN = true : integrator path, reversible kinematics (see above); cy = the work item's carry column
// LOCK (error-controlled kernel, LEAN only): the CTA's threads meet at every body; a thread with on = false just keeps pace.
SBK_HD void sweepBarrier() {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
}
template <bool LEAN, int JMASK = JM_ALL, bool CB = false, bool LOCK = false> SBK_HD void tpiEvalDerivatives(const Ctx& c, const Tables& T, int inst, double* cy, double* qdotDst, double* udotDst, double* qddDst, const bool on = true) {
    if constexpr (!LEAN) {
        tpiKinematics<CB>(c, inst, qdotDst);
        twoPointPass<CB>(c, inst);
        tpiInward<IN_ABI | IN_Z | IN_BIAS | IN_FORCES, CB>(c, inst);
        tpiOutward<true, CB>(c, inst, udotDst, qddDst);
    } else {
        const SV z0 = zeroSV();
        // Two coordinate / G-nu slots alternate from body step to body step.  The slot parity follows from the
        // body index alone (outward sweeps: (b-1)&1, inward sweep: b&1 -- consecutive across the sweep changes),
        // so no step counter has to live across the joint switch.
        #define SBK_PRE(par) (cy + (CY_PRE + 4*((par) & 1))*SBK_CARRY_STRIDE)
        #define SBK_GNU(par) (cy + (CY_GNU + GNU_ROWS*((par) & 1))*SBK_CARRY_STRIDE)
        volatile int* bslot = reinterpret_cast<volatile int*>(cy + CY_LOOP*SBK_CARRY_STRIDE);
        #define SBK_PARK(b)   (*bslot = (b))
        #define SBK_UNPARK(b) ((b) = *bslot)
        if (on) { cyStoreOut(cy, identity3(), zero3(), z0);                       // Ground's link for body 1
                  preloadCoords<JMASK>(c, inst, T.bodies[1], SBK_PRE(0)); }
#pragma unroll 1
        for (int b = 1; b < c.nb; ++b) {
            if constexpr (LOCK) sweepBarrier();
            if (!on) continue;
            const BodyConst& bc = T.bodies[b];
            preloadCoords<JMASK>(c, inst, T.bodies[b + 1 < c.nb ? b + 1 : c.nb - 1], SBK_PRE(b));       // after the last body: the first of the inward sweep
            preloadWait();
            SBK_PARK(b);
            SBK_DISPATCH_JOINT_M(JMASK, bc.joint, (leanKinBody<JT>(c, bc, inst, cy, SBK_PRE(b - 1), qdotDst)));
            SBK_UNPARK(b);
        }
#pragma unroll 1
        for (int b = c.nb - 1; b >= 1; --b) {
            if constexpr (LOCK) sweepBarrier();
            if (!on) continue;
            const BodyConst& bc = T.bodies[b];
            preloadCoords<JMASK>(c, inst, T.bodies[b > 1 ? b - 1 : 1], SBK_PRE(b - 1));                  // after body 1: the first of the outward sweep
            preloadWait();
            SBK_PARK(b);
            SBK_DISPATCH_JOINT_M(JMASK, bc.joint, (leanInwardBody<JT>(c, T, bc, b, inst, cy, SBK_PRE(b))));
            SBK_UNPARK(b);
        }
        if (on) { cyStoreOut(cy, identity3(), zero3(), z0); cyStoreA(cy + CY_A*SBK_CARRY_STRIDE, z0); }
#pragma unroll 1
        for (int b = 1; b < c.nb; ++b) {
            if constexpr (LOCK) sweepBarrier();
            if (!on) continue;
            const BodyConst& bc = T.bodies[b];
            if (b + 1 < c.nb) preloadGNu(c, inst, T.bodies[b + 1], SBK_GNU(b));
            preloadCoords<JMASK>(c, inst, T.bodies[b + 1 < c.nb ? b + 1 : b], SBK_PRE(b));
            preloadWait();
            // body 1 wrote its G / nu at the very end of the inward sweep: it loads them directly
            SBK_PARK(b);
            SBK_DISPATCH_JOINT_M(JMASK, bc.joint, (leanOutwardBody<JT>(c, bc, inst, cy, SBK_PRE(b - 1), b > 1 ? SBK_GNU(b - 1) : nullptr, udotDst, qddDst)));
            SBK_UNPARK(b);
        }
        #undef SBK_PRE
        #undef SBK_GNU
        #undef SBK_PARK
        #undef SBK_UNPARK
    }
}

} // namespace sbkd
