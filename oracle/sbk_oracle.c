/* sbk_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the reference's
 * algorithm for the hot path (simbody/simbody 3.9.0): one derivative evaluation of a
 * tree-topology system (sweeps A-E + gravity/spring/damper forces), the matter operators, and
 * the fixed-step Runge-Kutta-Merson step.  One instance at a time, array-of-structures, no
 * vectorisation: clarity over speed.  It is the checker for the CUDA path; nothing in the
 * product links, imports or executes it (only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline leg may).
 *
 * PARITY PINNING: this restatement is pinned against the UNMODIFIED reference itself
 * (oracle/_ref, compiled from /root/reference by oracle/Makefile) through the golden vectors
 * in tests/golden/*.npz and live differential runs (tests/test_oracle.py); SURVEY.md section 8c
 * golden values (udot of the README double pendulum and the mixed 7-body fixture) are checked
 * too.  The only third-party arithmetic on the path is LAPACK dgetrf/dgetri for the 6x6 D^-1 of
 * Free mobilizers (SmallMatrixMixed.h:859-893; any LAPACK >= 3.6, OpenBLAS 0.3.15 in
 * oracle/_ref); it is restated here as LU with partial pivoting, the published algorithm of
 * dgetrf + dgetri.
 *
 * Reference lines followed (paths relative to the reference root):
 *   slot rules            Simbody/src/RigidBodyNodeSpec.h:81-87
 *   X_FM, H_FM, HDot_FM   Simbody/src/RigidBodyNodeSpec_{Pin,Slider,Universal,Ball,Free}.h
 *   sweep A               Simbody/src/RigidBodyNodeSpec.h:229-288,554-569; RigidBodyNodeSpec.cpp:44-74;
 *                         RigidBodyNode.cpp:54-84
 *   sweep B               RigidBodyNodeSpec.h:305-333; RigidBodyNodeSpec.cpp:82-129; RigidBodyNode.cpp:97-174
 *   sweep C               RigidBodyNodeSpec.cpp:249-325; MassProperties.cpp:76-127
 *   bias                  RigidBodyNode.cpp:201-213
 *   forces                Force_Gravity.cpp:514-569; Force.cpp:339-351,434-443
 *   sweeps D, E           RigidBodyNodeSpec.cpp:355-446
 *   operators             RigidBodyNodeSpec.cpp:483-695
 *   RKM step              SimTKmath/Integrators/src/RungeKuttaMersonIntegrator.cpp:86-140,
 *                         AbstractIntegratorRep.cpp:137-208, IntegratorRep.h:454-513,644-655,
 *                         SimbodyMatterSubsystemRep.cpp:4096-4184,4448-4476
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { PIN = 1, SLIDER = 2, UNIVERSAL = 3, BALL = 4, FREE = 5, WELD = 6, TRANSLATION = 7, CYLINDER = 8, PLANAR = 9, GIMBAL = 10,
       BALL_EULER = 11, FREE_EULER = 12 };   /* Ball / Free under setUseEulerAngles: x-y-z angles, last q slot unused */
enum { F_GRAVITY = 1, F_SPRING = 2, F_DAMPER = 3, F_UNIFORM_GRAVITY = 4, F_GLOBAL_DAMPER = 5, F_MOBILITY_CONSTANT = 6,
       F_TWO_POINT_SPRING = 7, F_TWO_POINT_DAMPER = 8 };
#define MAXD 6

typedef struct {
    int nb, nf;
    const int *parent, *joint;
    const double *mass, *com, *uinertia, *X_PF, *X_BM;   /* [nb], [nb*3], [nb*6], [nb*12], [nb*12] */
    const int *fkind, *fbody, *fcoord; const double *fa, *fb, *fdir;   /* [nf], ..., [nf*6]: dir (3) then station2 (3) per force */
} Model;

typedef struct {   /* per-body cache, dense */
    double R[9], p[3];          /* X_GB */
    double V[6], A[6];          /* V_GB, A_GB (angular, linear) */
    double H[MAXD][6];          /* columns of H_PB_G */
    double l[3];                /* Phi */
    double m, c[3], G[9];       /* Mk_G: mass, com in G, unit inertia in G (full 3x3) */
    double a[6], b[6];          /* mobilizer coriolis acceleration, gyroscopic force */
    double P[36], PP[36];       /* articulated inertia and P+ as dense 6x6: [J F; F^T M] */
    double DI[MAXD*MAXD], Gm[MAXD][6];
    double zb[6], z[6], zP[6], eps[MAXD], Fapp[6];
    int q0, u0, nq, nu;
} Body;

static int NQ(int j) { return j == PIN || j == SLIDER ? 1 : (j == UNIVERSAL || j == CYLINDER) ? 2 : (j == TRANSLATION || j == PLANAR || j == GIMBAL) ? 3 : (j == BALL || j == BALL_EULER) ? 4 : (j == FREE || j == FREE_EULER) ? 7 : 0; }
static int NU(int j) { return j == PIN || j == SLIDER ? 1 : (j == UNIVERSAL || j == CYLINDER) ? 2 : (j == BALL || j == BALL_EULER || j == TRANSLATION || j == PLANAR || j == GIMBAL) ? 3 : (j == FREE || j == FREE_EULER) ? 6 : 0; }

static void matvec3(const double* R, const double* v, double* o) { for (int i = 0; i < 3; ++i) o[i] = R[3*i]*v[0] + R[3*i+1]*v[1] + R[3*i+2]*v[2]; }
static void matmul3(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3*i+j] = A[3*i]*B[j] + A[3*i+1]*B[3+j] + A[3*i+2]*B[6+j];
}
static void cross(const double* a, const double* b, double* o) {
    double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0]; o[0] = x; o[1] = y; o[2] = z;
}
static void crossMat(const double* v, double* M) { M[0]=0; M[1]=-v[2]; M[2]=v[1]; M[3]=v[2]; M[4]=0; M[5]=-v[0]; M[6]=-v[1]; M[7]=v[0]; M[8]=0; }
/* ~Phi(l)*V and Phi(l)*F (SpatialAlgebra.h:729-762) */
static void phiT(const double* l, const double* V, double* o) { double t[3]; cross(V, l, t); for (int i = 0; i < 3; ++i) { o[i] = V[i]; o[3+i] = V[3+i] + t[i]; } }
static void phiF(const double* l, const double* F, double* o) { double t[3]; cross(l, F+3, t); for (int i = 0; i < 3; ++i) { o[i] = F[i] + t[i]; o[3+i] = F[3+i]; } }
static void mat6vec(const double* P, const double* v, double* o) { for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += P[6*i+j]*v[j]; o[i] = s; } }
/* SpatialInertia * V (MassProperties.h:1071-1072) */
static void mkTimes(const Body* B, const double* V, double* o) {
    double Gw[3], pv[3], pw[3]; matvec3(B->G, V, Gw); cross(B->c, V+3, pv); cross(B->c, V, pw);
    for (int i = 0; i < 3; ++i) { o[i] = B->m*(Gw[i] + pv[i]); o[3+i] = B->m*(V[3+i] - pw[i]); }
}

/* General inverse by LU with partial pivoting (dgetrf + dgetri); n <= 6. Closed forms for n <= 3
 * as in SmallMatrixMixed.h:841-1006. */
static int invertD(int n, const double* D, double* DI) {
    if (n == 0) return 1;
    if (n == 1) { DI[0] = 1.0/D[0]; return D[0] != 0; }
    if (n == 2) { double det = D[0]*D[3] - D[1]*D[2], ood = 1.0/det; DI[0] = ood*D[3]; DI[1] = -ood*D[1]; DI[2] = -ood*D[2]; DI[3] = ood*D[0]; return det != 0; }
    if (n == 3) {
        #define m(i,j) D[3*(i)+(j)]
        double d00 = m(1,1)*m(2,2)-m(1,2)*m(2,1), nd01 = m(1,2)*m(2,0)-m(1,0)*m(2,2), d02 = m(1,0)*m(2,1)-m(1,1)*m(2,0);
        double det = m(0,0)*d00 + m(0,1)*nd01 + m(0,2)*d02, ood = 1.0/det;
        double nd10 = m(0,2)*m(2,1)-m(0,1)*m(2,2), d11 = m(0,0)*m(2,2)-m(0,2)*m(2,0), nd12 = m(0,1)*m(2,0)-m(0,0)*m(2,1),
               d20 = m(0,1)*m(1,2)-m(0,2)*m(1,1), nd21 = m(0,2)*m(1,0)-m(0,0)*m(1,2), d22 = m(0,0)*m(1,1)-m(0,1)*m(1,0);
        #undef m
        DI[0] = ood*d00; DI[1] = ood*nd10; DI[2] = ood*d20; DI[3] = ood*nd01; DI[4] = ood*d11; DI[5] = ood*nd21; DI[6] = ood*d02; DI[7] = ood*nd12; DI[8] = ood*d22;
        return det != 0;
    }
    double A[MAXD*MAXD]; int piv[MAXD];
    memcpy(A, D, sizeof(double)*n*n);
    for (int k = 0; k < n; ++k) {
        int p = k; double best = fabs(A[n*k+k]);
        for (int i = k+1; i < n; ++i) if (fabs(A[n*i+k]) > best) { best = fabs(A[n*i+k]); p = i; }
        piv[k] = p; if (best == 0) return 0;
        if (p != k) for (int j = 0; j < n; ++j) { double t = A[n*k+j]; A[n*k+j] = A[n*p+j]; A[n*p+j] = t; }
        for (int i = k+1; i < n; ++i) { A[n*i+k] /= A[n*k+k]; for (int j = k+1; j < n; ++j) A[n*i+j] -= A[n*i+k]*A[n*k+j]; }
    }
    for (int c = 0; c < n; ++c) {            /* solve A x = P e_c */
        double x[MAXD]; for (int i = 0; i < n; ++i) x[i] = (i == c);
        for (int k = 0; k < n; ++k) if (piv[k] != k) { double t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
        for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) x[i] -= A[n*i+j]*x[j];
        for (int i = n-1; i >= 0; --i) { for (int j = i+1; j < n; ++j) x[i] -= A[n*i+j]*x[j]; x[i] /= A[n*i+i]; }
        for (int i = 0; i < n; ++i) DI[n*i+c] = x[i];
    }
    return 1;
}

static void quatN(const double* q, const double* w, double* o) {   /* Rotation.h:712-720 */
    double e0 = q[0]/2, e1 = q[1]/2, e2 = q[2]/2, e3 = q[3]/2;
    o[0] = -e1*w[0] + -e2*w[1] + -e3*w[2]; o[1] = e0*w[0] + e3*w[1] + -e2*w[2];
    o[2] = -e3*w[0] + e0*w[1] + e1*w[2];   o[3] = e2*w[0] + -e1*w[1] + e0*w[2];
}
static void quatNInv(const double* q, const double* qd, double* o) {   /* Rotation.h:742-748 */
    double e0 = 2*q[0], e1 = 2*q[1], e2 = 2*q[2], e3 = 2*q[3];
    o[0] = -e1*qd[0] + e0*qd[1] + -e3*qd[2] + e2*qd[3];
    o[1] = -e2*qd[0] + e3*qd[1] + e0*qd[2] + -e1*qd[3];
    o[2] = -e3*qd[0] + -e2*qd[1] + e1*qd[2] + e0*qd[3];
}

static Body* setupBodies(const Model* M, int* nqOut, int* nuOut, int* nquatOut) {
    Body* B = (Body*)calloc((size_t)M->nb, sizeof(Body));
    int q = 0, u = 0, nquat = 0;
    for (int b = 1; b < M->nb; ++b) {
        B[b].q0 = q; B[b].u0 = u; B[b].nq = NQ(M->joint[b]); B[b].nu = NU(M->joint[b]); q += B[b].nq; u += B[b].nu;
        if (M->joint[b] == BALL || M->joint[b] == FREE) ++nquat;
    }
    B[0].R[0] = B[0].R[4] = B[0].R[8] = 1;
    *nqOut = q; *nuOut = u; *nquatOut = nquat;
    return B;
}

/* Sweeps A + B for all bodies, base to tip (body index order is a valid order). */
static void kinematics(const Model* M, Body* B, const double* q, const double* u, double* qdot, double* qerr) {
    int iq = 0;
    for (int b = 1; b < M->nb; ++b) {
        Body* me = &B[b]; const Body* pa = &B[M->parent[b]];
        const int jt = M->joint[b], d = me->nu; const double* qb = q + me->q0; const double* ub = u + me->u0;
        double Rfm[9] = {1,0,0, 0,1,0, 0,0,1}, pfm[3] = {0,0,0};
        double Hw[MAXD][3], Hv[MAXD][3], HDw[MAXD][3];
        memset(Hw, 0, sizeof Hw); memset(Hv, 0, sizeof Hv); memset(HDw, 0, sizeof HDw);
        if (jt == PIN) { double c = cos(qb[0]), s = sin(qb[0]); Rfm[0] = c; Rfm[1] = -s; Rfm[3] = s; Rfm[4] = c; Hw[0][2] = 1; }
        else if (jt == SLIDER) { pfm[0] = qb[0]; Hv[0][0] = 1; }
        else if (jt == UNIVERSAL) {   /* body-fixed X then Y (Rotation.cpp:241-264) */
            double c1 = cos(qb[0]), s1 = sin(qb[0]), c2 = cos(qb[1]), s2 = sin(qb[1]);
            Rfm[0] = c2; Rfm[1] = 0; Rfm[2] = s2; Rfm[3] = s2*s1; Rfm[4] = c1; Rfm[5] = -s1*c2; Rfm[6] = -s2*c1; Rfm[7] = s1; Rfm[8] = c1*c2;
            Hw[0][0] = 1; Hw[1][0] = Rfm[1]; Hw[1][1] = Rfm[4]; Hw[1][2] = Rfm[7];
        } else if (jt == WELD) {      /* X_FM = I, no mobilities (RigidBodyNode_Weld.cpp:369-420) */
        } else if (jt == TRANSLATION) { pfm[0] = qb[0]; pfm[1] = qb[1]; pfm[2] = qb[2]; Hv[0][0] = Hv[1][1] = Hv[2][2] = 1;   /* _Translation.h:100-130 */
        } else if (jt == CYLINDER) {  /* _Cylinder.h:110-139 */
            double c = cos(qb[0]), s = sin(qb[0]); Rfm[0] = c; Rfm[1] = -s; Rfm[3] = s; Rfm[4] = c; pfm[2] = qb[1]; Hw[0][2] = 1; Hv[1][2] = 1;
        } else if (jt == BALL_EULER || jt == FREE_EULER) {   /* _Ball.h:118-160, _Free.h:147-180: x-y-z angles, H_FM = I */
            double c0 = cos(qb[0]), c1 = cos(qb[1]), c2 = cos(qb[2]), s0 = sin(qb[0]), s1 = sin(qb[1]), s2 = sin(qb[2]);
            double s0s1 = s0*s1, s2c0 = s2*c0, c0c2 = c0*c2, nc1 = -c1;
            Rfm[0] = c1*c2; Rfm[1] = s2*nc1; Rfm[2] = s1; Rfm[3] = s2c0 + s0s1*c2; Rfm[4] = c0c2 - s0s1*s2; Rfm[5] = s0*nc1;
            Rfm[6] = s0*s2 - s1*c0c2; Rfm[7] = s0*c2 + s1*s2c0; Rfm[8] = c0*c1;
            Hw[0][0] = Hw[1][1] = Hw[2][2] = 1;
            if (jt == FREE_EULER) { pfm[0] = qb[3]; pfm[1] = qb[4]; pfm[2] = qb[5]; Hv[3][0] = Hv[4][1] = Hv[5][2] = 1; }
        } else if (jt == GIMBAL) {    /* _Gimbal.h:108-176, Rotation.h:342-349; u = qdot */
            double c0 = cos(qb[0]), c1 = cos(qb[1]), c2 = cos(qb[2]), s0 = sin(qb[0]), s1 = sin(qb[1]), s2 = sin(qb[2]);
            double s0s1 = s0*s1, s2c0 = s2*c0, c0c2 = c0*c2, nc1 = -c1;
            Rfm[0] = c1*c2; Rfm[1] = s2*nc1; Rfm[2] = s1; Rfm[3] = s2c0 + s0s1*c2; Rfm[4] = c0c2 - s0s1*s2; Rfm[5] = s0*nc1;
            Rfm[6] = s0*s2 - s1*c0c2; Rfm[7] = s0*c2 + s1*s2c0; Rfm[8] = c0*c1;
            Hw[0][0] = 1; Hw[1][1] = c0; Hw[1][2] = s0; Hw[2][0] = s1; Hw[2][1] = -s0*c1; Hw[2][2] = c0*c1;
            { double qd0 = ub[0], qd1 = ub[1], dc0 = -s0*qd0, dc1 = -s1*qd1, ds0 = c0*qd0, ds1 = c1*qd1;
              HDw[1][1] = dc0; HDw[1][2] = ds0; HDw[2][0] = ds1; HDw[2][1] = -ds0*c1 - s0*dc1; HDw[2][2] = dc0*c1 + c0*dc1; }
        } else if (jt == PLANAR) {    /* _Planar.h:126-157 */
            double c = cos(qb[0]), s = sin(qb[0]); Rfm[0] = c; Rfm[1] = -s; Rfm[3] = s; Rfm[4] = c; pfm[0] = qb[1]; pfm[1] = qb[2]; Hw[0][2] = 1; Hv[1][0] = 1; Hv[2][1] = 1;
        } else {                      /* Ball / Free, quaternion (Rotation.cpp:600-611) */
            double n = sqrt(qb[0]*qb[0] + qb[1]*qb[1] + qb[2]*qb[2] + qb[3]*qb[3]);
            if (qerr) qerr[iq] = n - 1.0; ++iq;
            double oo = 1.0/n, a0 = qb[0]*oo, a1 = qb[1]*oo, a2 = qb[2]*oo, a3 = qb[3]*oo;
            double q00=a0*a0, q11=a1*a1, q22=a2*a2, q33=a3*a3, q01=a0*a1, q02=a0*a2, q03=a0*a3, q12=a1*a2, q13=a1*a3, q23=a2*a3;
            double q00mq11 = q00-q11, q22mq33 = q22-q33;
            Rfm[0] = q00+q11-q22-q33; Rfm[1] = 2*(q12-q03); Rfm[2] = 2*(q13+q02);
            Rfm[3] = 2*(q12+q03); Rfm[4] = q00mq11+q22mq33; Rfm[5] = 2*(q23-q01);
            Rfm[6] = 2*(q13-q02); Rfm[7] = 2*(q23+q01); Rfm[8] = q00mq11-q22mq33;
            Hw[0][0] = Hw[1][1] = Hw[2][2] = 1;
            if (jt == FREE) { pfm[0] = qb[4]; pfm[1] = qb[5]; pfm[2] = qb[6]; Hv[3][0] = Hv[4][1] = Hv[5][2] = 1; }
        }
        /* X_MB = ~X_BM */
        const double* XBM = M->X_BM + 12*b; const double* XPF = M->X_PF + 12*b;
        double Rmb[9], pmb[3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rmb[3*i+j] = XBM[3*j+i];
        for (int i = 0; i < 3; ++i) pmb[i] = -(Rmb[3*i]*XBM[9] + Rmb[3*i+1]*XBM[10] + Rmb[3*i+2]*XBM[11]);
        /* X_FB = X_FM X_MB; X_PB = X_PF X_FB; X_GB = X_GP X_PB */
        double r[3], Rfb[9], pfb[3], Rpb[9], ppb[3], t[3];
        matvec3(Rfm, pmb, r); matmul3(Rfm, Rmb, Rfb); for (int i = 0; i < 3; ++i) pfb[i] = pfm[i] + r[i];
        matmul3(XPF, Rfb, Rpb); matvec3(XPF, pfb, t); for (int i = 0; i < 3; ++i) ppb[i] = XPF[9+i] + t[i];
        matmul3(pa->R, Rpb, me->R); matvec3(pa->R, ppb, me->l); for (int i = 0; i < 3; ++i) me->p[i] = pa->p[i] + me->l[i];
        /* H = R_GF (H_FM + H_MB_F) */
        double Rgf[9]; matmul3(pa->R, XPF, Rgf);
        for (int j = 0; j < d; ++j) {
            double hv[3], cx[3]; cross(Hw[j], r, cx); for (int i = 0; i < 3; ++i) hv[i] = Hv[j][i] + cx[i];
            matvec3(Rgf, Hw[j], me->H[j]); matvec3(Rgf, hv, me->H[j]+3);
        }
        /* mass properties in G: G_G = R G_B R^T (dense; the reference uses the 57-flop form) */
        const double* ui = M->uinertia + 6*b;
        double Gb[9] = {ui[0], ui[3], ui[4], ui[3], ui[1], ui[5], ui[4], ui[5], ui[2]}, RG[9], Rt[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[3*i+j] = me->R[3*j+i];
        matmul3(me->R, Gb, RG); matmul3(RG, Rt, me->G);
        me->m = M->mass[b]; matvec3(me->R, M->com + 3*b, me->c);
        /* velocity */
        double wfm[3] = {0,0,0}, Vpb[6] = {0,0,0,0,0,0};
        for (int j = 0; j < d; ++j) { for (int i = 0; i < 3; ++i) wfm[i] += Hw[j][i]*ub[j]; for (int i = 0; i < 6; ++i) Vpb[i] += me->H[j][i]*ub[j]; }
        if (jt == UNIVERSAL) { double y[3] = {Rfm[1], Rfm[4], Rfm[7]}; cross(wfm, y, HDw[1]); }
        double wxr[3]; cross(wfm, r, wxr);
        double VD[6] = {0,0,0,0,0,0};
        for (int j = 0; j < d; ++j) {
            double HD[6], t1[3], t2[3], t3[3], s[3];
            matvec3(Rgf, HDw[j], t1); cross(pa->V, me->H[j], t2); for (int i = 0; i < 3; ++i) HD[i] = t1[i] + t2[i];
            cross(HDw[j], r, t1); cross(Hw[j], wxr, t2); for (int i = 0; i < 3; ++i) s[i] = t1[i] + t2[i];
            matvec3(Rgf, s, t3); cross(pa->V, me->H[j]+3, t2); for (int i = 0; i < 3; ++i) HD[3+i] = t3[i] + t2[i];
            for (int i = 0; i < 6; ++i) VD[i] += HD[i]*ub[j];
        }
        double Vs[6]; phiT(me->l, pa->V, Vs); for (int i = 0; i < 6; ++i) me->V[i] = Vs[i] + Vpb[i];
        double Gw[3], wGw[3], wc[3], wwc[3], dv[3], wdv[3];
        matvec3(me->G, me->V, Gw); cross(me->V, Gw, wGw); cross(me->V, me->c, wc); cross(me->V, wc, wwc);
        for (int i = 0; i < 3; ++i) { me->b[i] = me->m*wGw[i]; me->b[3+i] = me->m*wwc[i]; }
        for (int i = 0; i < 3; ++i) dv[i] = me->V[3+i] - pa->V[3+i];
        cross(pa->V, dv, wdv);
        for (int i = 0; i < 3; ++i) { me->a[i] = VD[i]; me->a[3+i] = VD[3+i] + wdv[i]; }
        /* qdot */
        if (qdot) {
            if (jt == BALL_EULER || jt == FREE_EULER) {   /* qdot = N_P w (Rotation.h:395-406); unused slot 0 */
                double c0 = cos(qb[0]), s0 = sin(qb[0]), s1 = sin(qb[1]), oc = 1/cos(qb[1]), t = (s0*ub[1] - c0*ub[2])*oc;
                qdot[me->q0] = ub[0] + t*s1; qdot[me->q0+1] = c0*ub[1] + s0*ub[2]; qdot[me->q0+2] = -t;
                if (jt == FREE_EULER) { for (int i = 0; i < 3; ++i) qdot[me->q0+3+i] = ub[3+i]; qdot[me->q0+6] = 0; } else qdot[me->q0+3] = 0;
            }
            else if (jt == BALL || jt == FREE) { quatN(qb, ub, qdot + me->q0); if (jt == FREE) for (int i = 0; i < 3; ++i) qdot[me->q0+4+i] = ub[3+i]; }
            else for (int i = 0; i < d; ++i) qdot[me->q0+i] = ub[i];
        }
    }
}

/* Sweep C: dense 6x6 articulated inertias, tip to base. */
static int articulatedInertias(const Model* M, Body* B) {
    int ok = 1;
    for (int b = M->nb-1; b >= 1; --b) {
        Body* me = &B[b]; const int d = me->nu;
        double mc[3] = {me->m*me->c[0], me->m*me->c[1], me->m*me->c[2]}, F[9]; crossMat(mc, F);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            me->P[6*i+j] = me->m*me->G[3*i+j]; me->P[6*i+3+j] = F[3*i+j]; me->P[6*(3+i)+j] = F[3*j+i]; me->P[6*(3+i)+3+j] = (i == j) ? me->m : 0.0;
        }
        for (int c = b+1; c < M->nb; ++c) if (M->parent[c] == b) {   /* P += Phi(l) P+ ~Phi(l), dense */
            double Phi[36] = {0}, T[36], S[36], lx[9]; crossMat(B[c].l, lx);
            for (int i = 0; i < 6; ++i) Phi[7*i] = 1;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Phi[6*i+3+j] = lx[3*i+j];
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += Phi[6*i+k]*B[c].PP[6*k+j]; T[6*i+j] = s; }
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += T[6*i+k]*Phi[6*j+k]; S[6*i+j] = s; }
            for (int i = 0; i < 36; ++i) me->P[i] += S[i];
        }
        double PH[MAXD][6], D[MAXD*MAXD];
        for (int j = 0; j < d; ++j) mat6vec(me->P, me->H[j], PH[j]);
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += me->H[i][k]*PH[j][k]; D[d*i+j] = s; }
        if (!invertD(d, D, me->DI)) ok = 0;
        for (int j = 0; j < d; ++j) for (int k = 0; k < 6; ++k) { double s = 0; for (int i = 0; i < d; ++i) s += PH[i][k]*me->DI[d*i+j]; me->Gm[j][k] = s; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += me->Gm[k][i]*PH[k][j]; me->PP[6*i+j] = me->P[6*i+j] - s; }
        for (int i = 0; i < 6; ++i) for (int j = i+1; j < 6; ++j) { double s = 0.5*(me->PP[6*i+j] + me->PP[6*j+i]); me->PP[6*i+j] = me->PP[6*j+i] = s; }
        double Pa[6]; mat6vec(me->P, me->a, Pa); for (int i = 0; i < 6; ++i) me->zb[i] = Pa[i] + me->b[i];
    }
    return ok;
}

static void systemForces(const Model* M, Body* B, const double* q, const double* u, double* fmob) {
    int nu = 0; for (int b = 1; b < M->nb; ++b) nu += B[b].nu;
    for (int i = 0; i < nu; ++i) fmob[i] = 0;
    for (int b = 0; b < M->nb; ++b) memset(B[b].Fapp, 0, sizeof B[b].Fapp);
    for (int k = 0; k < M->nf; ++k) {
        if (M->fkind[k] == F_GRAVITY || M->fkind[k] == F_UNIFORM_GRAVITY) {   /* Force_Gravity.cpp:532 g*d; Force.cpp:1053 the vector itself */
            const int uni = M->fkind[k] == F_UNIFORM_GRAVITY;
            const double g[3] = {uni ? M->fdir[6*k] : M->fa[k]*M->fdir[6*k], uni ? M->fdir[6*k+1] : M->fa[k]*M->fdir[6*k+1], uni ? M->fdir[6*k+2] : M->fa[k]*M->fdir[6*k+2]};
            for (int b = 1; b < M->nb; ++b) { double F[3] = {B[b].m*g[0], B[b].m*g[1], B[b].m*g[2]}, t[3]; cross(B[b].c, F, t);
                for (int i = 0; i < 3; ++i) { B[b].Fapp[i] += t[i]; B[b].Fapp[3+i] += F[i]; } }
        } else if (M->fkind[k] == F_SPRING) { const Body* me = &B[M->fbody[k]]; fmob[me->u0 + M->fcoord[k]] += -M->fa[k]*(q[me->q0 + M->fcoord[k]] - M->fb[k]); }
        else if (M->fkind[k] == F_MOBILITY_CONSTANT) { const Body* me = &B[M->fbody[k]]; fmob[me->u0 + M->fcoord[k]] += M->fa[k]; }
        else if (M->fkind[k] == F_GLOBAL_DAMPER) { for (int i = 0; i < nu; ++i) fmob[i] -= M->fa[k]*u[i]; }   /* Force.cpp:997 */
        else if (M->fkind[k] == F_DAMPER) { const Body* me = &B[M->fbody[k]]; fmob[me->u0 + M->fcoord[k]] += -M->fa[k]*u[me->u0 + M->fcoord[k]]; }
        else if (M->fkind[k] == F_TWO_POINT_SPRING || M->fkind[k] == F_TWO_POINT_DAMPER) {   /* Force.cpp:103-140, 179-221 */
            Body* b1 = &B[M->fbody[k]]; Body* b2 = &B[M->fcoord[k]];
            double s1[3], s2[3], r[3], f1[3], t1[3], t2[3];
            matvec3(b1->R, M->fdir + 6*k, s1); matvec3(b2->R, M->fdir + 6*k + 3, s2);
            for (int i = 0; i < 3; ++i) r[i] = (b2->p[i] + s2[i]) - (b1->p[i] + s1[i]);
            const double d = sqrt(r[0]*r[0] + r[1]*r[1] + r[2]*r[2]);
            if (M->fkind[k] == F_TWO_POINT_SPRING) { const double frc = M->fa[k]*(d - M->fb[k]); for (int i = 0; i < 3; ++i) f1[i] = (frc/d)*r[i]; }
            else {
                double w1[3], w2[3], dir[3], vr = 0;
                cross(b1->V, s1, w1); cross(b2->V, s2, w2);
                for (int i = 0; i < 3; ++i) { dir[i] = r[i]/d; vr += ((b2->V[3+i] + w2[i]) - (b1->V[3+i] + w1[i]))*dir[i]; }
                for (int i = 0; i < 3; ++i) f1[i] = M->fa[k]*vr*dir[i];
            }
            cross(s1, f1, t1); cross(s2, f1, t2);
            for (int i = 0; i < 3; ++i) { b1->Fapp[i] += t1[i]; b1->Fapp[3+i] += f1[i]; b2->Fapp[i] -= t2[i]; b2->Fapp[3+i] -= f1[i]; }
        }
    }
}

/* Sweeps D and E. withBias: z starts from P a + b - F (forward dynamics) or 0 (M^-1). */
static void accelerations(const Model* M, Body* B, const double* fmob, int withBias, double* udot) {
    for (int b = M->nb-1; b >= 1; --b) {
        Body* me = &B[b]; const int d = me->nu;
        for (int i = 0; i < 6; ++i) me->z[i] = withBias ? me->zb[i] - me->Fapp[i] : 0.0;
        for (int c = b+1; c < M->nb; ++c) if (M->parent[c] == b) { double t[6]; phiF(B[c].l, B[c].zP, t); for (int i = 0; i < 6; ++i) me->z[i] += t[i]; }
        for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += me->H[j][k]*me->z[k]; me->eps[j] = fmob[me->u0+j] - s; }
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < d; ++j) s += me->Gm[j][i]*me->eps[j]; me->zP[i] = me->z[i] + s; }
    }
    memset(B[0].A, 0, sizeof B[0].A);
    for (int b = 1; b < M->nb; ++b) {
        Body* me = &B[b]; const int d = me->nu; double Ap[6];
        phiT(me->l, B[M->parent[b]].A, Ap);
        for (int i = 0; i < d; ++i) { double s = 0, g = 0; for (int j = 0; j < d; ++j) s += me->DI[d*i+j]*me->eps[j]; for (int k = 0; k < 6; ++k) g += me->Gm[i][k]*Ap[k]; udot[me->u0+i] = s - g; }
        for (int k = 0; k < 6; ++k) { double s = 0; for (int j = 0; j < d; ++j) s += me->H[j][k]*udot[me->u0+j]; me->A[k] = Ap[k] + s + (withBias ? me->a[k] : 0.0); }
    }
}

static void qdotdot(const Model* M, const Body* B, const double* q, const double* u, const double* udot, double* qdd) {
    for (int b = 1; b < M->nb; ++b) {
        const Body* me = &B[b]; const int jt = M->joint[b];
        if (jt == BALL_EULER || jt == FREE_EULER) {   /* Rotation.h:1041-1060 */
            const double* qb = q + me->q0; const double* w = u + me->u0; const double* b = udot + me->u0;
            double c0 = cos(qb[0]), s0 = sin(qb[0]), c1 = cos(qb[1]), s1 = sin(qb[1]), oc = 1/c1;
            double t = (s0*w[1] - c0*w[2])*oc, qd0 = w[0] + t*s1, qd1 = c0*w[1] + s0*w[2], qd2 = -t;
            double tb = (s0*b[1] - c0*b[2])*oc, q1oc1 = qd1*oc;
            qdd[me->q0] = (b[0] + tb*s1) + (qd0*s1 - qd2)*q1oc1; qdd[me->q0+1] = (c0*b[1] + s0*b[2]) + qd0*qd2*c1; qdd[me->q0+2] = -tb + (qd2*s1 - qd0)*q1oc1;
            if (jt == FREE_EULER) { for (int i = 0; i < 3; ++i) qdd[me->q0+3+i] = udot[me->u0+3+i]; qdd[me->q0+6] = 0; } else qdd[me->q0+3] = 0;
        } else if (jt == BALL || jt == FREE) {
            const double* w = u + me->u0; double Nb[4]; quatN(q + me->q0, udot + me->u0, Nb);
            double k = -0.25*(w[0]*w[0] + w[1]*w[1] + w[2]*w[2]);
            for (int i = 0; i < 4; ++i) qdd[me->q0+i] = Nb[i] + k*q[me->q0+i];
            if (jt == FREE) for (int i = 0; i < 3; ++i) qdd[me->q0+4+i] = udot[me->u0+3+i];
        } else for (int i = 0; i < me->nu; ++i) qdd[me->q0+i] = udot[me->u0+i];
    }
}

/* M*v (withVel=0) or inverse dynamics residual (withVel=1). */
static void inverseDynamics(const Model* M, Body* B, const double* udotIn, const double* fmob, const double* Fbody, int withVel, double* tau) {
    memset(B[0].A, 0, sizeof B[0].A);
    for (int b = 1; b < M->nb; ++b) {
        Body* me = &B[b]; double Ap[6]; phiT(me->l, B[M->parent[b]].A, Ap);
        for (int k = 0; k < 6; ++k) { double s = 0; for (int j = 0; j < me->nu; ++j) s += me->H[j][k]*(udotIn ? udotIn[me->u0+j] : 0.0); me->A[k] = Ap[k] + s + (withVel ? me->a[k] : 0.0); }
    }
    for (int b = M->nb-1; b >= 1; --b) {
        Body* me = &B[b]; double F[6]; mkTimes(me, me->A, F);
        if (withVel) for (int k = 0; k < 6; ++k) F[k] = F[k] + me->b[k] - (Fbody ? Fbody[6*b+k] : 0.0);
        for (int c = b+1; c < M->nb; ++c) if (M->parent[c] == b) { double t[6]; phiF(B[c].l, B[c].zP, t); for (int k = 0; k < 6; ++k) F[k] += t[k]; }
        memcpy(me->zP, F, sizeof F);
        for (int j = 0; j < me->nu; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += me->H[j][k]*F[k]; tau[me->u0+j] = s - ((withVel && fmob) ? fmob[me->u0+j] : 0.0); }
    }
}

static void derivs(const Model* M, Body* B, int nq, int nu, const double* y, double* ydot, double* qdd, double* qerr, double* fmobOut) {
    double* fm = (double*)malloc(sizeof(double)*(size_t)(nu > 0 ? nu : 1));
    kinematics(M, B, y, y+nq, ydot, qerr);
    systemForces(M, B, y, y+nq, fm);
    articulatedInertias(M, B);
    accelerations(M, B, fm, 1, ydot+nq);
    if (qdd) qdotdot(M, B, y, y+nq, ydot+nq, qdd);
    if (fmobOut) memcpy(fmobOut, fm, sizeof(double)*(size_t)nu);
    free(fm);
}

static Model mkModel(int nb, const int* parent, const int* joint, const double* mass, const double* com, const double* ui,
                     const double* XPF, const double* XBM, int nf, const int* fkind, const int* fbody, const int* fcoord,
                     const double* fa, const double* fb, const double* fdir) {
    Model M; M.nb = nb; M.nf = nf; M.parent = parent; M.joint = joint; M.mass = mass; M.com = com; M.uinertia = ui; M.X_PF = XPF; M.X_BM = XBM;
    M.fkind = fkind; M.fbody = fbody; M.fcoord = fcoord; M.fa = fa; M.fb = fb; M.fdir = fdir; return M;
}

/* Same instance-major binary layout as `ref_driver eval` (oracle/ref_driver.cpp). */
int oracle_eval(int nb, const int* parent, const int* joint, const double* mass, const double* com, const double* ui,
                const double* XPF, const double* XBM, int nf, const int* fkind, const int* fbody, const int* fcoord,
                const double* fa, const double* fb, const double* fdir, int N, const double* in, double* out) {
    Model M = mkModel(nb, parent, joint, mass, com, ui, XPF, XBM, nf, fkind, fbody, fcoord, fa, fb, fdir);
    int nq, nu, nquat; Body* B = setupBodies(&M, &nq, &nu, &nquat);
    const int inStride = nq + 5*nu + 6*nb;
    double* ydot = (double*)malloc(sizeof(double)*(size_t)(nq+nu+1)); double* tmp = (double*)malloc(sizeof(double)*(size_t)(nu+1));
    for (int k = 0; k < N; ++k) {
        const double* p = in + (size_t)k*inStride; double* o = out;
        const int outStride = nq + nu + nq + nquat + nb*12 + nb*6 + nb*6 + nu + nb*6 + 4*nu + nu + nb*6;
        o = out + (size_t)k*outStride;
        double* qdd = o + nq + nu; double* qerr = qdd + nq; double* X = qerr + nquat; double* V = X + nb*12; double* A = V + nb*6;
        double* fsys = A + nb*6; double* Fsys = fsys + nu; double* Ma = Fsys + nb*6; double* MInv = Ma + nu; double* res = MInv + nu;
        double* res0 = res + nu; double* udop = res0 + nu; double* Aop = udop + nu;
        derivs(&M, B, nq, nu, p, ydot, qdd, qerr, fsys);
        memcpy(o, ydot, sizeof(double)*(size_t)(nq+nu));
        for (int b = 0; b < nb; ++b) { memcpy(X + 12*b, B[b].R, 9*sizeof(double)); memcpy(X + 12*b + 9, B[b].p, 3*sizeof(double));
                                       memcpy(V + 6*b, B[b].V, 6*sizeof(double)); memcpy(A + 6*b, B[b].A, 6*sizeof(double)); memcpy(Fsys + 6*b, B[b].Fapp, 6*sizeof(double)); }
        const double* pa = p + nq + nu; const double* pv = pa + nu; const double* pud = pv + nu; const double* pf = pud + nu; const double* pF = pf + nu;
        inverseDynamics(&M, B, pa, NULL, NULL, 0, Ma);
        accelerations(&M, B, pv, 0, MInv);
        inverseDynamics(&M, B, pud, pf, pF, 1, res);
        inverseDynamics(&M, B, NULL, NULL, NULL, 1, res0);
        for (int b = 0; b < nb; ++b) memcpy(B[b].Fapp, pF + 6*b, 6*sizeof(double));
        accelerations(&M, B, pf, 1, udop);
        for (int b = 0; b < nb; ++b) memcpy(Aop + 6*b, B[b].A, 6*sizeof(double));
        (void)tmp;
    }
    free(ydot); free(tmp); free(B);
    return 0;
}

/* calcMobilizerReactionForces (SimbodyMatterSubsystemRep.cpp:5788-5832): FB = zPlus + PPlus*(~Phi A_GP), reported at
 * the M frame origin: FM = (m - (R_GB p_BM) x f, f).  Ground collects its base bodies (RigidBodyNode_Weld.cpp:197-222).
 * multiplyBySystemJacobian / Transpose: RigidBodyNodeSpec.cpp:760-815.
 * calcCompositeBodyInertias: RigidBodyNode.cpp:231-243, MassProperties.h:1037-1045,1110-1116 (SpatialInertia += and shift).
 * in  per instance: q[nq] u[nu] v[nu] F[nb*6];  out: FM_G[nb*6] Jv[nb*6] JtF[nu] CBI[nb*10]   (same layout as `ref_driver extras`) */
int oracle_extras(int nb, const int* parent, const int* joint, const double* mass, const double* com, const double* ui,
                  const double* XPF, const double* XBM, int nf, const int* fkind, const int* fbody, const int* fcoord,
                  const double* fa, const double* fb, const double* fdir, int N, const double* in, double* out) {
    Model M = mkModel(nb, parent, joint, mass, com, ui, XPF, XBM, nf, fkind, fbody, fcoord, fa, fb, fdir);
    int nq, nu, nquat; Body* B = setupBodies(&M, &nq, &nu, &nquat);
    const int inStride = nq + 2*nu + 6*nb, outStride = 12*nb + nu + 10*nb;
    double* ydot = (double*)malloc(sizeof(double)*(size_t)(nq+nu+1));
    double* qerr = (double*)malloc(sizeof(double)*(size_t)(nquat+1));
    double* z = (double*)malloc(sizeof(double)*(size_t)(6*nb));
    for (int k = 0; k < N; ++k) {
        const double* p = in + (size_t)k*inStride; double* o = out + (size_t)k*outStride;
        derivs(&M, B, nq, nu, p, ydot, NULL, qerr, NULL);
        /* reactions */
        double z0[6]; for (int i = 0; i < 6; ++i) z0[i] = -B[0].Fapp[i];
        for (int c = 1; c < nb; ++c) if (parent[c] == 0) { double t[6]; phiF(B[c].l, B[c].zP, t); for (int i = 0; i < 6; ++i) z0[i] += t[i]; }
        for (int i = 0; i < 6; ++i) o[i] = z0[i];
        for (int b = 1; b < nb; ++b) {
            const Body* me = &B[b]; const Body* pa = &B[parent[b]];
            double pPB[3] = {me->p[0]-pa->p[0], me->p[1]-pa->p[1], me->p[2]-pa->p[2]}, Ap[6], PA[6], FB[6], pBM[3], t[3];
            phiT(pPB, pa->A, Ap); mat6vec(me->PP, Ap, PA);
            for (int i = 0; i < 6; ++i) FB[i] = me->zP[i] + PA[i];
            matvec3(me->R, M.X_BM + 12*b + 9, pBM); cross(pBM, FB+3, t);
            for (int i = 0; i < 3; ++i) { o[6*b+i] = FB[i] - t[i]; o[6*b+3+i] = FB[3+i]; }
        }
        /* J v */
        const double* v = p + nq + nu; double* Jv = o + 6*nb;
        for (int i = 0; i < 6; ++i) Jv[i] = 0;
        for (int b = 1; b < nb; ++b) {
            const Body* me = &B[b]; double sh[6]; phiT(me->l, Jv + 6*parent[b], sh);
            for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < me->nu; ++j) s += me->H[j][i]*v[me->u0+j]; Jv[6*b+i] = sh[i] + s; }
        }
        /* ~J F */
        const double* F = p + nq + 2*nu; double* JtF = o + 12*nb;
        memcpy(z, F, sizeof(double)*(size_t)(6*nb));
        for (int b = nb-1; b >= 1; --b) {
            const Body* me = &B[b];
            for (int c = b+1; c < nb; ++c) if (parent[c] == b) { double t[6]; phiF(B[c].l, z + 6*c, t); for (int i = 0; i < 6; ++i) z[6*b+i] += t[i]; }
            for (int j = 0; j < me->nu; ++j) { double s = 0; for (int i = 0; i < 6; ++i) s += me->H[j][i]*z[6*b+i]; JtF[me->u0+j] = s; }
        }
        /* composite body inertias, tip to base: R_b = Mk_b + sum R_c.shift(-l_c) */
        double* R = o + 12*nb + nu;
        R[0] = INFINITY; R[1] = R[2] = R[3] = 0; R[4] = R[5] = R[6] = 1; R[7] = R[8] = R[9] = 0;
        for (int b = nb-1; b >= 1; --b) {
            const Body* me = &B[b]; double* Rb = R + 10*b;
            Rb[0] = me->m; for (int i = 0; i < 3; ++i) Rb[1+i] = me->c[i];
            Rb[4] = me->G[0]; Rb[5] = me->G[4]; Rb[6] = me->G[8]; Rb[7] = me->G[1]; Rb[8] = me->G[2]; Rb[9] = me->G[5];
            for (int c = b+1; c < nb; ++c) if (parent[c] == b) {
                const double* Rc = R + 10*c; double p[3] = {Rc[1], Rc[2], Rc[3]}, G[6] = {Rc[4], Rc[5], Rc[6], Rc[7], Rc[8], Rc[9]}, pn[3];
                /* shift(S = -l): to centroid, then to the new origin: p' = p - S = p + l */
                G[0] -= p[1]*p[1]+p[2]*p[2]; G[1] -= p[0]*p[0]+p[2]*p[2]; G[2] -= p[0]*p[0]+p[1]*p[1]; G[3] -= -p[0]*p[1]; G[4] -= -p[0]*p[2]; G[5] -= -p[1]*p[2];
                for (int i = 0; i < 3; ++i) pn[i] = p[i] + B[c].l[i];
                G[0] += pn[1]*pn[1]+pn[2]*pn[2]; G[1] += pn[0]*pn[0]+pn[2]*pn[2]; G[2] += pn[0]*pn[0]+pn[1]*pn[1]; G[3] += -pn[0]*pn[1]; G[4] += -pn[0]*pn[2]; G[5] += -pn[1]*pn[2];
                /* += */
                const double mt = Rb[0] + Rc[0], oo = 1.0/mt;
                for (int i = 0; i < 3; ++i) Rb[1+i] = oo*(Rb[0]*Rb[1+i] + Rc[0]*pn[i]);
                for (int i = 0; i < 6; ++i) Rb[4+i] = oo*(Rb[0]*Rb[4+i] + Rc[0]*G[i]);
                Rb[0] = mt;
            }
        }
    }
    free(ydot); free(qerr); free(z); free(B);
    return 0;
}

static double errNorm(const Model* M, const Body* B, int nq, int nu, const double* y1, const double* y0, const double* err, int infNorm) {
    double qa = 0, ua = 0;
    for (int i = 0; i < nu; ++i) { double a = fabs(y0[nq+i]), sc = a > 1.0 ? 1.0/a : 1.0, v = sc*err[nq+i]; if (infNorm) { if (fabs(v) > ua) ua = fabs(v); } else ua += v*v; }
    for (int b = 1; b < M->nb; ++b) {
        const Body* me = &B[b]; int first = 0;
        if (M->joint[b] == BALL || M->joint[b] == FREE) { double du[3], o[4]; quatNInv(y1+me->q0, err+me->q0, du); quatN(y1+me->q0, du, o);
            for (int i = 0; i < 4; ++i) { if (infNorm) { if (fabs(o[i]) > qa) qa = fabs(o[i]); } else qa += o[i]*o[i]; } first = 4; }
        for (int i = first; i < me->nq; ++i) { double v = err[me->q0+i]; if (infNorm) { if (fabs(v) > qa) qa = fabs(v); } else qa += v*v; }
    }
    double qn = infNorm ? qa : (nq ? sqrt(qa/nq) : 0), un = infNorm ? ua : (nu ? sqrt(ua/nu) : 0);
    return qn >= un ? qn : un;
}

/* in [N][ny]; out [N][ny+2]: y after nsteps, error norm of the last step, number of projections */
int oracle_step(int nb, const int* parent, const int* joint, const double* mass, const double* com, const double* ui,
                const double* XPF, const double* XBM, int nf, const int* fkind, const int* fbody, const int* fcoord,
                const double* fa, const double* fb, const double* fdir, int N, const double* in, double* out,
                double h, int nsteps, double accuracy, double consTol, int infNorm, int projectEveryStep) {
    Model M = mkModel(nb, parent, joint, mass, com, ui, XPF, XBM, nf, fkind, fbody, fcoord, fa, fb, fdir);
    int nq, nu, nquat; Body* B = setupBodies(&M, &nq, &nu, &nquat);
    const int ny = nq + nu;
    double* w = (double*)malloc(sizeof(double)*(size_t)ny*7);
    double *y = w, *y0 = w+ny, *f0 = w+2*ny, *fa_ = w+3*ny, *fb_ = w+4*ny, *ys = w+5*ny, *er = w+6*ny;
    for (int k = 0; k < N; ++k) {
        memcpy(y, in + (size_t)k*ny, sizeof(double)*(size_t)ny);
        double en = 0; int nproj = 0;
        for (int s = 0; s < nsteps; ++s) {
            derivs(&M, B, nq, nu, y, f0, NULL, NULL, NULL);
            memcpy(y0, y, sizeof(double)*(size_t)ny);
            for (int i = 0; i < ny; ++i) y[i] = y0[i] + (h/3)*f0[i];
            derivs(&M, B, nq, nu, y, fa_, NULL, NULL, NULL);
            for (int i = 0; i < ny; ++i) y[i] = y0[i] + (h/6)*(f0[i] + fa_[i]);
            derivs(&M, B, nq, nu, y, fa_, NULL, NULL, NULL);
            for (int i = 0; i < ny; ++i) y[i] = y0[i] + (h/8)*(f0[i] + 3*fa_[i]);
            derivs(&M, B, nq, nu, y, fb_, NULL, NULL, NULL);
            for (int i = 0; i < ny; ++i) { ys[i] = y0[i] + (h/2)*(f0[i] - 3*fa_[i] + 4*fb_[i]); y[i] = ys[i]; }
            derivs(&M, B, nq, nu, y, fa_, NULL, NULL, NULL);
            for (int i = 0; i < ny; ++i) { y[i] = y0[i] + (h/6)*(f0[i] + 4*fb_[i] + fa_[i]); er[i] = 0.2*fabs(y[i] - ys[i]); }
            en = errNorm(&M, B, nq, nu, y, y0, er, infNorm);
            if (nquat > 0 && !(en > 16.0*accuracy)) {
                double acc = 0;
                for (int b = 1; b < nb; ++b) if (joint[b] == BALL || joint[b] == FREE) { const double* qq = y + B[b].q0;
                    double e = sqrt(qq[0]*qq[0] + qq[1]*qq[1] + qq[2]*qq[2] + qq[3]*qq[3]) - 1.0; if (infNorm) { if (fabs(e) > acc) acc = fabs(e); } else acc += e*e; }
                double qn = infNorm ? acc : sqrt(acc/nquat);
                /* AbstractIntegratorRep.cpp:165-190: beyond max(2 tol, sqrt(tol)) the step is a convergence failure: error norm = Infinity, no projection */
                const double plim = (2*consTol > sqrt(consTol)) ? 2*consTol : sqrt(consTol);
                if (qn > plim) en = INFINITY;
                else if (qn > consTol || projectEveryStep) {
                    for (int b = 1; b < nb; ++b) if (joint[b] == BALL || joint[b] == FREE) { double* qq = y + B[b].q0; double* ee = er + B[b].q0;
                        double n = sqrt(qq[0]*qq[0] + qq[1]*qq[1] + qq[2]*qq[2] + qq[3]*qq[3]), dt = 0;
                        for (int i = 0; i < 4; ++i) { qq[i] = qq[i]/n; dt += ee[i]*qq[i]; }
                        for (int i = 0; i < 4; ++i) ee[i] -= dt*qq[i]; }
                    ++nproj; en = errNorm(&M, B, nq, nu, y, y0, er, infNorm);
                }
            }
        }
        memcpy(out + (size_t)k*(ny+2), y, sizeof(double)*(size_t)ny);
        out[(size_t)k*(ny+2)+ny] = en; out[(size_t)k*(ny+2)+ny+1] = nproj;
    }
    free(w); free(B);
    return 0;
}
