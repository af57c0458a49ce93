// sbk_math.cuh -- fixed-size FP64 vector / spatial algebra used by every sweep.
//
// Everything here is a plain inline function on register-resident structs so that, after
// unrolling, a body step holds its 6x6 spatial algebra in registers (no local memory).
// Formulas follow the reference's spatial-algebra conventions:
//   SpatialVec = (angular, linear)            SimTKcommon/Mechanics/.../SpatialAlgebra.h:62-108
//   Phi*v, ~Phi*v                             SpatialAlgebra.h:729-762
//   SpatialInertia * SpatialVec               MassProperties.h:1071-1072
//   ArticulatedInertia (M sym, J sym, F full) MassProperties.h:1235-1330, MassProperties.cpp:76-137
//   Rotation from quaternion                  Rotation.cpp:600-611
//   reexpress unit inertia                    Rotation.cpp:780-803
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SBK_HD  __host__ __device__ __forceinline__
// Per-body steps are compiled as separate device functions (one register allocation each):
// inlining all five mobilizers x all sweeps into one kernel body exhausts the 255-register
// budget and spills.
#define SBK_HDN __host__ __device__ __noinline__
#else
#define SBK_HD  inline
#define SBK_HDN inline
#endif

namespace sbkd {

struct V3 { double x, y, z; };
struct M3 { double a[9]; };                       // row-major
struct S3 { double xx, yy, zz, xy, xz, yz; };    // symmetric 3x3
struct SV { V3 w, v; };                           // spatial vector (angular, linear)
// Articulated-body inertia: P = [J F; ~F M]
struct ABI { S3 M; S3 J; M3 F; };

SBK_HD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SBK_HD V3 operator+(V3 a, V3 b) { return mk(a.x+b.x, a.y+b.y, a.z+b.z); }
SBK_HD V3 operator-(V3 a, V3 b) { return mk(a.x-b.x, a.y-b.y, a.z-b.z); }
SBK_HD V3 operator-(V3 a)       { return mk(-a.x, -a.y, -a.z); }
SBK_HD V3 operator*(double s, V3 a) { return mk(s*a.x, s*a.y, s*a.z); }
SBK_HD double dot(V3 a, V3 b)   { return a.x*b.x + a.y*b.y + a.z*b.z; }
SBK_HD V3 cross(V3 a, V3 b)     { return mk(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
SBK_HD V3 zero3()               { return mk(0, 0, 0); }

SBK_HD SV operator+(SV a, SV b) { SV r; r.w = a.w + b.w; r.v = a.v + b.v; return r; }
SBK_HD SV operator-(SV a, SV b) { SV r; r.w = a.w - b.w; r.v = a.v - b.v; return r; }
SBK_HD SV operator*(double s, SV a) { SV r; r.w = s*a.w; r.v = s*a.v; return r; }
SBK_HD SV zeroSV() { SV r; r.w = zero3(); r.v = zero3(); return r; }
SBK_HD double dot(SV a, SV b) { return dot(a.w, b.w) + dot(a.v, b.v); }

SBK_HD V3 mul(const M3& R, V3 v) {
    return mk(R.a[0]*v.x + R.a[1]*v.y + R.a[2]*v.z,
              R.a[3]*v.x + R.a[4]*v.y + R.a[5]*v.z,
              R.a[6]*v.x + R.a[7]*v.y + R.a[8]*v.z);
}
SBK_HD V3 mulT(const M3& R, V3 v) {   // ~R * v
    return mk(R.a[0]*v.x + R.a[3]*v.y + R.a[6]*v.z,
              R.a[1]*v.x + R.a[4]*v.y + R.a[7]*v.z,
              R.a[2]*v.x + R.a[5]*v.y + R.a[8]*v.z);
}
SBK_HD M3 mul(const M3& A, const M3& B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C.a[3*i+j] = A.a[3*i]*B.a[j] + A.a[3*i+1]*B.a[3+j] + A.a[3*i+2]*B.a[6+j];
    return C;
}
SBK_HD M3 mulABt(const M3& A, const M3& B) {   // A * ~B
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C.a[3*i+j] = A.a[3*i]*B.a[3*j] + A.a[3*i+1]*B.a[3*j+1] + A.a[3*i+2]*B.a[3*j+2];
    return C;
}
SBK_HD V3 col(const M3& R, int j) { return mk(R.a[j], R.a[3+j], R.a[6+j]); }
SBK_HD M3 identity3() { M3 R; R.a[0]=1; R.a[1]=0; R.a[2]=0; R.a[3]=0; R.a[4]=1; R.a[5]=0; R.a[6]=0; R.a[7]=0; R.a[8]=1; return R; }

SBK_HD V3 mul(const S3& S, V3 v) {
    return mk(S.xx*v.x + S.xy*v.y + S.xz*v.z,
              S.xy*v.x + S.yy*v.y + S.yz*v.z,
              S.xz*v.x + S.yz*v.y + S.zz*v.z);
}

// Rotation from a NORMALISED quaternion, scalar first (Rotation.cpp:600-611).
SBK_HD M3 rotFromQuat(double q0, double q1, double q2, double q3) {
    const double q00=q0*q0, q11=q1*q1, q22=q2*q2, q33=q3*q3;
    const double q01=q0*q1, q02=q0*q2, q03=q0*q3;
    const double q12=q1*q2, q13=q1*q3, q23=q2*q3;
    const double q00mq11 = q00-q11, q22mq33 = q22-q33;
    M3 R;
    R.a[0] = q00+q11-q22-q33; R.a[1] = 2*(q12-q03);     R.a[2] = 2*(q13+q02);
    R.a[3] = 2*(q12+q03);     R.a[4] = q00mq11+q22mq33; R.a[5] = 2*(q23-q01);
    R.a[6] = 2*(q13-q02);     R.a[7] = 2*(q23+q01);     R.a[8] = q00mq11-q22mq33;
    return R;
}

// G_G = R_GB * G_B * ~R_GB for a symmetric G_B.  RigidBodyNode.cpp:76 calls
// UnitInertia::reexpress(~R_GB), which evaluates (~~R_GB).reexpressSymMat33(G_B)
// (MassProperties.h:808-809), i.e. Rotation_::reexpressSymMat33 (Rotation.cpp:780-803) on R_GB.
SBK_HD S3 reexpressSym(const M3& Rgb, const S3& S) {
    #define RR(i,j) Rgb.a[3*(i)+(j)]
    const double a=S.xx, b=S.yy, c=S.zz, d=S.xy, e=S.xz, f=S.yz;
    // L = [a-c d; d b-c; 2e 2f]   (3x2)
    const double L00=a-c, L01=d, L10=d, L11=b-c, L20=2*e, L21=2*f;
    // Y = [R[1]*L(0) R[1]*L(1); R[2]*L(0) R[2]*L(1)], R[i] = row i, L(j) = column j
    const double Y00 = RR(1,0)*L00 + RR(1,1)*L10 + RR(1,2)*L20;
    const double Y01 = RR(1,0)*L01 + RR(1,1)*L11 + RR(1,2)*L21;
    const double Y10 = RR(2,0)*L00 + RR(2,1)*L10 + RR(2,2)*L20;
    const double Y11 = RR(2,0)*L01 + RR(2,1)*L11 + RR(2,2)*L21;
    // RR_ = first two columns of R; Zij = Y[i-1] * ~RR_[j]
    const double Z10 = Y00*RR(0,0) + Y01*RR(0,1), Z11 = Y00*RR(1,0) + Y01*RR(1,1);
    const double Z20 = Y10*RR(0,0) + Y11*RR(0,1), Z21 = Y10*RR(1,0) + Y11*RR(1,1),
                 Z22 = Y10*RR(2,0) + Y11*RR(2,1);
    const double Z00 = (L00+L11) - (Z11+Z22);
    const double Rv0 = RR(0,1)*e - RR(0,0)*f, Rv1 = RR(1,1)*e - RR(1,0)*f, Rv2 = RR(2,1)*e - RR(2,0)*f;
    S3 out;
    out.xx = Z00 + c;
    out.xy = Z10 + Rv2; out.yy = Z11 + c;
    out.xz = Z20 - Rv1; out.yz = Z21 + Rv0; out.zz = Z22 + c;
    #undef RR
    return out;
}

// ~Phi(l) * V : shift a velocity/acceleration outward (SpatialAlgebra.h:729-762)
SBK_HD SV phiT(V3 l, SV V) { SV r; r.w = V.w; r.v = V.v + cross(V.w, l); return r; }
// Phi(l) * F : shift a force inward
SBK_HD SV phi(V3 l, SV F)  { SV r; r.w = F.w + cross(l, F.v); r.v = F.v; return r; }

// SpatialInertia(m, p, G) * V  (MassProperties.h:1071-1072)
SBK_HD SV mulSpatialInertia(double m, V3 p, const S3& G, SV V) {
    SV r; r.w = m*(mul(G, V.w) + cross(p, V.v)); r.v = m*(V.v - cross(p, V.w)); return r;
}

// ArticulatedInertia from a rigid-body spatial inertia (MassProperties.h:1256-1257):
// M = m*I, J = m*G, F = crossMat(m*p)
SBK_HD ABI abiFromRigid(double m, V3 p, const S3& G) {
    ABI P;
    P.M.xx = m; P.M.yy = m; P.M.zz = m; P.M.xy = 0; P.M.xz = 0; P.M.yz = 0;
    P.J.xx = m*G.xx; P.J.yy = m*G.yy; P.J.zz = m*G.zz; P.J.xy = m*G.xy; P.J.xz = m*G.xz; P.J.yz = m*G.yz;
    const V3 mp = m*p;
    P.F.a[0] = 0;     P.F.a[1] = -mp.z; P.F.a[2] = mp.y;
    P.F.a[3] = mp.z;  P.F.a[4] = 0;     P.F.a[5] = -mp.x;
    P.F.a[6] = -mp.y; P.F.a[7] = mp.x;  P.F.a[8] = 0;
    return P;
}
// P * V  (MassProperties.h:1285-1286): (J w + F v, ~F w + M v)
SBK_HD SV mul(const ABI& P, SV V) {
    SV r; r.w = mul(P.J, V.w) + mul(P.F, V.v); r.v = mulT(P.F, V.w) + mul(P.M, V.v); return r;
}
// ArticulatedInertia::shift(s) (MassProperties.cpp:101-127) -- note: shifts by -s.
//   F' = F + s x M ;  J' = J + halfCrossDiff(s, ~F, F')
SBK_HD ABI shiftABI(const ABI& P, V3 s) {
    ABI R; R.M = P.M;
    // s % M for symmetric M: column j of result = s x (column j of M)
    const V3 m0 = mk(P.M.xx, P.M.xy, P.M.xz), m1 = mk(P.M.xy, P.M.yy, P.M.yz), m2 = mk(P.M.xz, P.M.yz, P.M.zz);
    const V3 c0 = cross(s, m0), c1 = cross(s, m1), c2 = cross(s, m2);
    M3 Fp;
    Fp.a[0] = P.F.a[0] + c0.x; Fp.a[1] = P.F.a[1] + c1.x; Fp.a[2] = P.F.a[2] + c2.x;
    Fp.a[3] = P.F.a[3] + c0.y; Fp.a[4] = P.F.a[4] + c1.y; Fp.a[5] = P.F.a[5] + c2.y;
    Fp.a[6] = P.F.a[6] + c0.z; Fp.a[7] = P.F.a[7] + c1.z; Fp.a[8] = P.F.a[8] + c2.z;
    R.F = Fp;
    // halfCrossDiff(v=s, A=~F, G=F')  (MassProperties.cpp:105-114); A(i,j) = F(j,i)
    #define A_(i,j) P.F.a[3*(j)+(i)]
    #define G_(i,j) Fp.a[3*(i)+(j)]
    const double v0 = s.x, v1 = s.y, v2 = s.z;
    R.J.xx = P.J.xx + (v1*(A_(2,0)+G_(0,2)) - v2*(A_(1,0)+G_(0,1)));
    R.J.xy = P.J.xy + (v2*(A_(0,0)-G_(1,1)) - v0*A_(2,0) + v1*G_(1,2));
    R.J.yy = P.J.yy + (v2*(A_(0,1)+G_(1,0)) - v0*(A_(2,1)+G_(1,2)));
    R.J.xz = P.J.xz + (v0*A_(1,0) - v2*G_(2,1) - v1*(A_(0,0)-G_(2,2)));
    R.J.yz = P.J.yz + (v0*(A_(1,1)-G_(2,2)) - v1*A_(0,1) + v2*G_(2,0));
    R.J.zz = P.J.zz + (v0*(A_(1,2)+G_(2,1)) - v1*(A_(0,2)+G_(2,0)));
    #undef A_
    #undef G_
    return R;
}
SBK_HD void addInto(ABI& P, const ABI& Q) {
    P.M.xx += Q.M.xx; P.M.yy += Q.M.yy; P.M.zz += Q.M.zz; P.M.xy += Q.M.xy; P.M.xz += Q.M.xz; P.M.yz += Q.M.yz;
    P.J.xx += Q.J.xx; P.J.yy += Q.J.yy; P.J.zz += Q.J.zz; P.J.xy += Q.J.xy; P.J.xz += Q.J.xz; P.J.yz += Q.J.yz;
#pragma unroll
    for (int i = 0; i < 9; ++i) P.F.a[i] += Q.F.a[i];
}

// ---- small dense inverses (SmallMatrixMixed.h:841-1006) -----------------------------------
// D is stored row-major d x d.  1x1 divide, 2x2 adjugate, 3x3 cofactors exactly as the
// reference; 6x6 (Free) is LAPACK getrf/getri in the reference -- here an unrolled LDL^T
// (D is symmetric positive definite), see DESIGN.md "6x6 inverse".
template <int d> struct Inv;
template <> struct Inv<1> {
    SBK_HD static bool run(const double* D, double* DI) { DI[0] = 1.0/D[0]; return D[0] != 0.0; }
};
template <> struct Inv<2> {
    SBK_HD static bool run(const double* D, double* DI) {
        const double det = D[0]*D[3] - D[1]*D[2];
        const double ood = 1.0/det;
        DI[0] = ood*D[3]; DI[1] = -ood*D[1]; DI[2] = -ood*D[2]; DI[3] = ood*D[0];
        return det != 0.0;
    }
};
template <> struct Inv<3> {
    SBK_HD static bool run(const double* m, double* DI) {
        #define m_(i,j) m[3*(i)+(j)]
        const double d00  = m_(1,1)*m_(2,2)-m_(1,2)*m_(2,1),
                     nd01 = m_(1,2)*m_(2,0)-m_(1,0)*m_(2,2),
                     d02  = m_(1,0)*m_(2,1)-m_(1,1)*m_(2,0);
        const double det = m_(0,0)*d00 + m_(0,1)*nd01 + m_(0,2)*d02;
        const double ood = 1.0/det;
        const double nd10 = m_(0,2)*m_(2,1)-m_(0,1)*m_(2,2),
                     d11  = m_(0,0)*m_(2,2)-m_(0,2)*m_(2,0),
                     nd12 = m_(0,1)*m_(2,0)-m_(0,0)*m_(2,1),
                     d20  = m_(0,1)*m_(1,2)-m_(0,2)*m_(1,1),
                     nd21 = m_(0,2)*m_(1,0)-m_(0,0)*m_(1,2),
                     d22  = m_(0,0)*m_(1,1)-m_(0,1)*m_(1,0);
        #undef m_
        DI[0] = ood*d00;  DI[1] = ood*nd10; DI[2] = ood*d20;
        DI[3] = ood*nd01; DI[4] = ood*d11;  DI[5] = ood*nd21;
        DI[6] = ood*d02;  DI[7] = ood*nd12; DI[8] = ood*d22;
        return det != 0.0;
    }
};
template <> struct Inv<6> {
    // Symmetric positive definite 6x6: D = L diag(e) L^T, then D^-1 = L^-T diag(1/e) L^-1.
    // Fully unrolled so every index is a compile-time constant (registers only).
    SBK_HD static bool run(const double* D, double* DI) {
        double L[36]; double e[6]; bool ok = true;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double s = D[7*j];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < j) s -= L[6*j+k]*L[6*j+k]*e[k];
            e[j] = s; ok = ok && (s > 0.0);
            const double oo = 1.0/s;
#pragma unroll
            for (int i = 0; i < 6; ++i) if (i > j) {
                double t = 0.5*(D[6*i+j] + D[6*j+i]);
#pragma unroll
                for (int k = 0; k < 6; ++k) if (k < j) t -= L[6*i+k]*L[6*j+k]*e[k];
                L[6*i+j] = t*oo;
            }
        }
        // W = L^-1 (unit lower triangular)
        double W[36];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
#pragma unroll
            for (int i = 0; i < 6; ++i) if (i > j) {
                double t = -L[6*i+j];
#pragma unroll
                for (int k = 0; k < 6; ++k) if (k > j && k < i) t -= L[6*i+k]*W[6*k+j];
                W[6*i+j] = t;
            }
        }
        // DI = W^T diag(1/e) W
        double oe[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) oe[k] = 1.0/e[k];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) if (j <= i) {
                // sum over k >= i of W[k][i]*W[k][j]/e[k], with W[k][k] = 1
                double t = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) if (k >= i) {
                    const double wki = (k == i) ? 1.0 : W[6*k+i];
                    const double wkj = (k == j) ? 1.0 : W[6*k+j];
                    t += wki*wkj*oe[k];
                }
                DI[6*i+j] = t; DI[6*j+i] = t;
            }
        }
        return ok;
    }
};

} // namespace sbkd
