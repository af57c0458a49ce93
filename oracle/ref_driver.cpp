// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Simbody-side harness, linked against the
// UNMODIFIED reference compiled by oracle/Makefile into oracle/_ref/libsimbody_ref.so.
//
// It builds a real SimTK::MultibodySystem from a model text (simbody_b200/host/model_spec.h)
// through the public Simbody API, and
//   lower  : prints the model lowered back from the realized system (lower_simbody.h)
//   slots  : prints the q/u slot map Simbody assigned
//   eval   : for N states, dumps realize(Acceleration) results and the matter operators
//            (calcAcceleration, multiplyByM, multiplyByMInv, calcResidualForceIgnoringConstraints)
//   extras : calcMobilizerReactionForces, multiplyBySystemJacobian[Transpose]
//   step   : RungeKuttaMersonIntegrator, fixed step h, nsteps steps, for N states
//   adaptive: RungeKuttaMersonIntegrator with error control to a final time (config C1)
//   bench  : CPU baseline -- one independent System+State+Integrator per host thread
//            (BASELINE.md section 3), prints one JSON line
// Binary I/O is raw little-endian float64, instance-major.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <thread>
#include <vector>

#include "Simbody.h"
#include "simbody_b200/host/model_spec.h"
#include "simbody_b200/host/lower_simbody.h"

using namespace SimTK;

static std::string slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "ref_driver: cannot open %s\n", path); std::exit(2); }
    std::ostringstream ss; ss << f.rdbuf(); return ss.str();
}
static std::vector<double> readDoubles(const char* path) {
    std::string s = slurp(path);
    std::vector<double> v(s.size()/sizeof(double));
    std::memcpy(v.data(), s.data(), v.size()*sizeof(double));
    return v;
}
static void writeDoubles(const char* path, const std::vector<double>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size()*sizeof(double)));
}
static Transform toTransform(const double X[12]) {
    Mat33 R(X[0],X[1],X[2], X[3],X[4],X[5], X[6],X[7],X[8]);
    return Transform(Rotation(R, true), Vec3(X[9],X[10],X[11]));
}

// A Simbody system built from a ModelSpec via the public API only.
struct RefSystem {
    MultibodySystem        system;
    SimbodyMatterSubsystem matter;
    GeneralForceSubsystem  forces;
    int nb=0, nq=0, nu=0, nquat=0;
    State defaultState;

    explicit RefSystem(const sbk::ModelSpec& spec) : matter(system), forces(system) {
        nb = (int)spec.bodies.size();
        for (int i = 1; i < nb; ++i) {
            const sbk_body_desc& b = spec.bodies[i];
            const double* ui = b.unit_inertia_OB_B;
            Body::Rigid body(MassProperties(b.mass, Vec3(b.com_B[0],b.com_B[1],b.com_B[2]),
                                            UnitInertia(ui[0],ui[1],ui[2],ui[3],ui[4],ui[5])));
            MobilizedBody& parent = matter.updMobilizedBody(MobilizedBodyIndex(b.parent));
            const Transform X_PF = toTransform(b.X_PF), X_BM = toTransform(b.X_BM);
            switch (b.joint_type) {
              case SBK_JOINT_PIN:       { MobilizedBody::Pin       m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_SLIDER:    { MobilizedBody::Slider    m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_UNIVERSAL: { MobilizedBody::Universal m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_BALL:      { MobilizedBody::Ball      m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_FREE:      { MobilizedBody::Free      m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_WELD:      { MobilizedBody::Weld      m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_TRANSLATION: { MobilizedBody::Translation m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_CYLINDER:  { MobilizedBody::Cylinder  m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_PLANAR:    { MobilizedBody::Planar    m(parent, X_PF, body, X_BM); break; }
              case SBK_JOINT_GIMBAL:    { MobilizedBody::Gimbal    m(parent, X_PF, body, X_BM); break; }
              default: throw std::runtime_error("ref_driver: bad joint type");
            }
        }
        for (const sbk_force_desc& f : spec.forces) {
            if (f.kind == SBK_FORCE_GRAVITY)
                Force::Gravity(forces, matter, UnitVec3(Vec3(f.dir[0],f.dir[1],f.dir[2])), f.a);
            else if (f.kind == SBK_FORCE_SPRING)
                Force::MobilityLinearSpring(forces, matter.getMobilizedBody(MobilizedBodyIndex(f.body)),
                                            MobilizerQIndex(f.coord), f.a, f.b);
            else if (f.kind == SBK_FORCE_DAMPER)
                Force::MobilityLinearDamper(forces, matter.getMobilizedBody(MobilizedBodyIndex(f.body)),
                                            MobilizerUIndex(f.coord), f.a);
            else if (f.kind == SBK_FORCE_UNIFORM_GRAVITY)
                Force::UniformGravity(forces, matter, Vec3(f.dir[0],f.dir[1],f.dir[2]));
            else if (f.kind == SBK_FORCE_GLOBAL_DAMPER)
                Force::GlobalDamper(forces, matter, f.a);
            else if (f.kind == SBK_FORCE_MOBILITY_CONSTANT)
                Force::MobilityConstantForce(forces, matter.getMobilizedBody(MobilizedBodyIndex(f.body)), MobilizerUIndex(f.coord), f.a);
            else if (f.kind == SBK_FORCE_TWO_POINT_SPRING)
                Force::TwoPointLinearSpring(forces, matter.getMobilizedBody(MobilizedBodyIndex(f.body)), Vec3(f.dir[0],f.dir[1],f.dir[2]),
                                            matter.getMobilizedBody(MobilizedBodyIndex(f.coord)), Vec3(f.station2[0],f.station2[1],f.station2[2]), f.a, f.b);
            else if (f.kind == SBK_FORCE_TWO_POINT_DAMPER)
                Force::TwoPointLinearDamper(forces, matter.getMobilizedBody(MobilizedBodyIndex(f.body)), Vec3(f.dir[0],f.dir[1],f.dir[2]),
                                            matter.getMobilizedBody(MobilizedBodyIndex(f.coord)), Vec3(f.station2[0],f.station2[1],f.station2[2]), f.a);
        }
        defaultState = system.realizeTopology();
        if (spec.useEulerAngles) matter.setUseEulerAngles(defaultState, true);
        system.realizeModel(defaultState);
        nq = defaultState.getNQ(); nu = defaultState.getNU();
        nquat = matter.getNumQuaternionsInUse(defaultState);
    }
};

static void setQU(State& s, const double* q, int nq, const double* u, int nu) {
    Vector& Q = s.updQ(); for (int i = 0; i < nq; ++i) Q[i] = q[i];
    Vector& U = s.updU(); for (int i = 0; i < nu; ++i) U[i] = u[i];
}

// ---- eval -----------------------------------------------------------------------------------
// in  per instance: q[nq] u[nu] a[nu] v[nu] known_udot[nu] fmob[nu] Fbody[nb*6]
// out per instance: qdot[nq] udot[nu] qdotdot[nq] qerr[nquat] X_GB[nb*12] V_GB[nb*6] A_GB[nb*6]
//                   fmob_sys[nu] Fbody_sys[nb*6] Ma[nu] MInvv[nu] resid[nu] resid0[nu]
//                   udot_op[nu] A_GB_op[nb*6]
static int cmdEval(RefSystem& rs, const char* inPath, const char* outPath, int N) {
    const int nb=rs.nb, nq=rs.nq, nu=rs.nu, nquat=rs.nquat;
    const int inStride  = nq + 5*nu + 6*nb;
    const int outStride = nq + nu + nq + nquat + nb*12 + nb*6 + nb*6 + nu + nb*6 + 4*nu + nu + nb*6;
    std::vector<double> in = readDoubles(inPath);
    if ((int)in.size() != N*inStride) { std::fprintf(stderr, "eval: input has %zu doubles, expected %d\n", in.size(), N*inStride); return 2; }
    std::vector<double> out((size_t)N*outStride);
    State s = rs.defaultState;
    for (int k = 0; k < N; ++k) {
        const double* p = &in[(size_t)k*inStride];
        double* o = &out[(size_t)k*outStride];
        setQU(s, p, nq, p+nq, nu);
        rs.system.realize(s, Stage::Acceleration);
        for (int i = 0; i < nq; ++i) *o++ = s.getQDot()[i];
        for (int i = 0; i < nu; ++i) *o++ = s.getUDot()[i];
        for (int i = 0; i < nq; ++i) *o++ = s.getQDotDot()[i];
        for (int i = 0; i < nquat; ++i) *o++ = s.getQErr()[s.getNQErr()-nquat+i];
        for (MobilizedBodyIndex b(0); b < nb; ++b) {
            const Transform& X = rs.matter.getMobilizedBody(b).getBodyTransform(s);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) *o++ = X.R()[i][j];
            for (int i = 0; i < 3; ++i) *o++ = X.p()[i];
        }
        for (MobilizedBodyIndex b(0); b < nb; ++b) {
            const SpatialVec& V = rs.matter.getMobilizedBody(b).getBodyVelocity(s);
            for (int i = 0; i < 3; ++i) *o++ = V[0][i]; for (int i = 0; i < 3; ++i) *o++ = V[1][i];
        }
        for (MobilizedBodyIndex b(0); b < nb; ++b) {
            const SpatialVec& A = rs.matter.getMobilizedBody(b).getBodyAcceleration(s);
            for (int i = 0; i < 3; ++i) *o++ = A[0][i]; for (int i = 0; i < 3; ++i) *o++ = A[1][i];
        }
        const Vector& fm = rs.system.getMobilityForces(s, Stage::Dynamics);
        const Vector_<SpatialVec>& Fb = rs.system.getRigidBodyForces(s, Stage::Dynamics);
        for (int i = 0; i < nu; ++i) *o++ = fm[i];
        for (int b = 0; b < nb; ++b) { for (int i = 0; i < 3; ++i) *o++ = Fb[b][0][i]; for (int i = 0; i < 3; ++i) *o++ = Fb[b][1][i]; }

        const double* pa = p + nq + nu; const double* pv = pa + nu; const double* pud = pv + nu;
        const double* pf = pud + nu;    const double* pF = pf + nu;
        Vector a(nu), v(nu), ud(nu), f(nu), res;
        Vector_<SpatialVec> F(nb);
        for (int i = 0; i < nu; ++i) { a[i]=pa[i]; v[i]=pv[i]; ud[i]=pud[i]; f[i]=pf[i]; }
        for (int b = 0; b < nb; ++b) F[b] = SpatialVec(Vec3(pF[6*b],pF[6*b+1],pF[6*b+2]), Vec3(pF[6*b+3],pF[6*b+4],pF[6*b+5]));
        rs.matter.multiplyByM(s, a, res);      for (int i = 0; i < nu; ++i) *o++ = res[i];
        rs.matter.multiplyByMInv(s, v, res);   for (int i = 0; i < nu; ++i) *o++ = res[i];
        rs.matter.calcResidualForceIgnoringConstraints(s, f, F, ud, res);
        for (int i = 0; i < nu; ++i) *o++ = res[i];
        rs.matter.calcResidualForceIgnoringConstraints(s, Vector(), Vector_<SpatialVec>(), Vector(), res);
        for (int i = 0; i < nu; ++i) *o++ = res[i];
        Vector udotOp; Vector_<SpatialVec> AOp;
        rs.matter.calcAcceleration(s, f, F, udotOp, AOp);
        for (int i = 0; i < nu; ++i) *o++ = udotOp[i];
        for (int b = 0; b < nb; ++b) { for (int i = 0; i < 3; ++i) *o++ = AOp[b][0][i]; for (int i = 0; i < 3; ++i) *o++ = AOp[b][1][i]; }
        if (o - &out[(size_t)k*outStride] != outStride) { std::fprintf(stderr, "eval: stride bug\n"); return 3; }
    }
    writeDoubles(outPath, out);
    return 0;
}

// ---- energy -----------------------------------------------------------------------------------
// in per instance: q[nq] u[nu]; out per instance: kinetic, potential
static int cmdEnergy(RefSystem& rs, const char* inPath, const char* outPath, int N) {
    const int nq=rs.nq, nu=rs.nu, ny=nq+nu;
    std::vector<double> in = readDoubles(inPath);
    if ((int)in.size() != N*ny) { std::fprintf(stderr, "energy: bad input size\n"); return 2; }
    std::vector<double> out((size_t)N*2);
    State s = rs.defaultState;
    for (int k = 0; k < N; ++k) {
        setQU(s, &in[(size_t)k*ny], nq, &in[(size_t)k*ny+nq], nu);
        rs.system.realize(s, Stage::Dynamics);
        out[2*k] = rs.system.calcKineticEnergy(s); out[2*k+1] = rs.system.calcPotentialEnergy(s);
    }
    writeDoubles(outPath, out);
    return 0;
}

// ---- extras ---------------------------------------------------------------------------------
// in  per instance: q[nq] u[nu] v[nu] F[nb*6]
// out per instance: FM_G[nb*6]  calcMobilizerReactionForces at realize(Acceleration)
//                   Jv[nb*6]    multiplyBySystemJacobian(v)
//                   JtF[nu]     multiplyBySystemJacobianTranspose(F)
//                   CBI[nb*10]  calcCompositeBodyInertias: mass, com(3), unit inertia xx yy zz xy xz yz
static int cmdExtras(RefSystem& rs, const char* inPath, const char* outPath, int N) {
    const int nb=rs.nb, nq=rs.nq, nu=rs.nu;
    const int inStride = nq + 2*nu + 6*nb, outStride = 12*nb + nu + 10*nb;
    std::vector<double> in = readDoubles(inPath);
    if ((int)in.size() != N*inStride) { std::fprintf(stderr, "extras: bad input size\n"); return 2; }
    std::vector<double> out((size_t)N*outStride);
    State s = rs.defaultState;
    for (int k = 0; k < N; ++k) {
        const double* p = &in[(size_t)k*inStride];
        double* o = &out[(size_t)k*outStride];
        setQU(s, p, nq, p+nq, nu);
        rs.system.realize(s, Stage::Acceleration);
        Vector_<SpatialVec> FM; rs.matter.calcMobilizerReactionForces(s, FM);
        for (int b = 0; b < nb; ++b) { for (int i = 0; i < 3; ++i) *o++ = FM[b][0][i]; for (int i = 0; i < 3; ++i) *o++ = FM[b][1][i]; }
        Vector v(nu); for (int i = 0; i < nu; ++i) v[i] = p[nq+nu+i];
        Vector_<SpatialVec> Jv; rs.matter.multiplyBySystemJacobian(s, v, Jv);
        for (int b = 0; b < nb; ++b) { for (int i = 0; i < 3; ++i) *o++ = Jv[b][0][i]; for (int i = 0; i < 3; ++i) *o++ = Jv[b][1][i]; }
        Vector_<SpatialVec> F(nb); const double* pF = p + nq + 2*nu;
        for (int b = 0; b < nb; ++b) F[b] = SpatialVec(Vec3(pF[6*b],pF[6*b+1],pF[6*b+2]), Vec3(pF[6*b+3],pF[6*b+4],pF[6*b+5]));
        Vector JtF; rs.matter.multiplyBySystemJacobianTranspose(s, F, JtF);
        for (int i = 0; i < nu; ++i) *o++ = JtF[i];
        Array_<SpatialInertia, MobilizedBodyIndex> R; rs.matter.calcCompositeBodyInertias(s, R);
        for (MobilizedBodyIndex b(0); b < nb; ++b) {
            *o++ = R[b].getMass(); for (int i = 0; i < 3; ++i) *o++ = R[b].getMassCenter()[i];
            const Vec3& mo = R[b].getUnitInertia().getMoments(); const Vec3& pr = R[b].getUnitInertia().getProducts();
            for (int i = 0; i < 3; ++i) *o++ = mo[i]; for (int i = 0; i < 3; ++i) *o++ = pr[i];
        }
    }
    writeDoubles(outPath, out);
    return 0;
}

// ---- step -----------------------------------------------------------------------------------
// in per instance: q[nq] u[nu]; out per instance: q[nq] u[nu] stepsTaken realizations qProjections
// Integrator options a command line may set (Integrator.h:352-394): constraint tolerance, infinity norm, project every step
struct IntegOpts { double consTol = -1; int infNorm = 0; int projectEvery = 0; };
static void applyOpts(RungeKuttaMersonIntegrator& integ, const IntegOpts& o) {
    if (o.consTol > 0) integ.setConstraintTolerance(o.consTol);
    if (o.infNorm) integ.setUseInfinityNorm(true);
    if (o.projectEvery) integ.setProjectEveryStep(true);
}
static void configureFixed(RungeKuttaMersonIntegrator& integ, double h, double accuracy, const IntegOpts& o = IntegOpts()) {
    integ.setFixedStepSize(h);
    integ.setAllowInterpolation(false);
    if (accuracy > 0) integ.setAccuracy(accuracy);
    applyOpts(integ, o);
}
// The first stepTo() after initialize() returns at once (StartOfContinuousInterval), so loop
// until the advanced time reaches tFinal; every internal step has size h.
static void advanceTo(RungeKuttaMersonIntegrator& integ, double tFinal) {
    while (integ.getAdvancedTime() < tFinal*(1 - 1e-12)) integ.stepTo(tFinal);
}
static int cmdStep(RefSystem& rs, const char* inPath, const char* outPath, int N, double h, int nsteps, double accuracy, const IntegOpts& opts) {
    const int nq=rs.nq, nu=rs.nu, ny=nq+nu;
    std::vector<double> in = readDoubles(inPath);
    if ((int)in.size() != N*ny) { std::fprintf(stderr, "step: input has %zu doubles, expected %d\n", in.size(), N*ny); return 2; }
    std::vector<double> out((size_t)N*(ny+3));
    for (int k = 0; k < N; ++k) {
        State s = rs.defaultState;
        setQU(s, &in[(size_t)k*ny], nq, &in[(size_t)k*ny+nq], nu);
        RungeKuttaMersonIntegrator integ(rs.system);
        configureFixed(integ, h, accuracy, opts);
        integ.initialize(s);
        advanceTo(integ, nsteps*h);
        const State& a = integ.getAdvancedState();
        double* o = &out[(size_t)k*(ny+3)];
        for (int i = 0; i < nq; ++i) *o++ = a.getQ()[i];
        for (int i = 0; i < nu; ++i) *o++ = a.getU()[i];
        *o++ = integ.getNumStepsTaken(); *o++ = integ.getNumRealizations(); *o++ = integ.getNumQProjections();
    }
    writeDoubles(outPath, out);
    return 0;
}

// ---- adaptive (config C1) -------------------------------------------------------------------
static int cmdAdaptive(RefSystem& rs, const char* inPath, const char* outPath, int N, double tFinal, double accuracy, bool allowInterpolation,
                       const IntegOpts& opts) {
    const int nq=rs.nq, nu=rs.nu, ny=nq+nu;
    std::vector<double> in = readDoubles(inPath);
    // out per instance: advanced q, u | steps taken | steps attempted | realizations | last step | advanced time | q projections
    std::vector<double> out((size_t)N*(ny+6));
    for (int k = 0; k < N; ++k) {
        State s = rs.defaultState;
        setQU(s, &in[(size_t)k*ny], nq, &in[(size_t)k*ny+nq], nu);
        RungeKuttaMersonIntegrator integ(rs.system);
        if (accuracy > 0) integ.setAccuracy(accuracy);
        if (!allowInterpolation) integ.setAllowInterpolation(false);
        applyOpts(integ, opts);
        TimeStepper ts(rs.system, integ);
        ts.initialize(s);
        ts.stepTo(tFinal);
        const State& a = integ.getAdvancedState();
        double* o = &out[(size_t)k*(ny+6)];
        for (int i = 0; i < nq; ++i) *o++ = a.getQ()[i];
        for (int i = 0; i < nu; ++i) *o++ = a.getU()[i];
        *o++ = integ.getNumStepsTaken(); *o++ = integ.getNumStepsAttempted();
        *o++ = integ.getNumRealizations(); *o++ = integ.getPreviousStepSizeTaken(); *o++ = integ.getAdvancedTime();
        *o++ = integ.getNumQProjections();
    }
    writeDoubles(outPath, out);
    return 0;
}

// ---- bench: the reference CPU path, one system per host thread ------------------------------
static int cmdBench(const sbk::ModelSpec& spec, const char* inPath, int N, double h, int nsteps, int nthreads, const char* outPath) {
    std::vector<double> in = readDoubles(inPath);
    std::vector<std::unique_ptr<RefSystem>> systems;
    for (int t = 0; t < nthreads; ++t) systems.emplace_back(new RefSystem(spec));
    const int nq = systems[0]->nq, nu = systems[0]->nu, ny = nq+nu;
    if ((int)in.size() != N*ny) { std::fprintf(stderr, "bench: input has %zu doubles, expected %d\n", in.size(), N*ny); return 2; }
    std::vector<double> out((size_t)N*ny);
    // Only the stepping is timed: State copy and Integrator::initialize are
    // set-up.  The reported rate is total instance-steps / the slowest thread's stepping time.
    std::vector<double> stepSeconds(nthreads, 0.0);
    auto worker = [&](int t) {
        RefSystem& rs = *systems[t];
        for (int k = t; k < N; k += nthreads) {
            State s = rs.defaultState;
            setQU(s, &in[(size_t)k*ny], nq, &in[(size_t)k*ny+nq], nu);
            RungeKuttaMersonIntegrator integ(rs.system);
            configureFixed(integ, h, -1);
            integ.initialize(s);
            const auto t0 = std::chrono::steady_clock::now();
            advanceTo(integ, nsteps*h);
            stepSeconds[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const State& a = integ.getAdvancedState();
            for (int i = 0; i < nq; ++i) out[(size_t)k*ny+i] = a.getQ()[i];
            for (int i = 0; i < nu; ++i) out[(size_t)k*ny+nq+i] = a.getU()[i];
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double sec = 0; for (double v : stepSeconds) sec = std::max(sec, v);
    if (outPath) writeDoubles(outPath, out);
    std::printf("{\"instance_steps_per_s\": %.6g, \"seconds\": %.6g, \"wall_seconds\": %.6g, \"instances\": %d, \"steps\": %d, \"threads\": %d, \"h\": %.17g}\n",
                (double)N*nsteps/sec, sec, wall, N, nsteps, nthreads, h);
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc < 3) {
            std::fprintf(stderr,
                "usage: ref_driver lower|slots <model.txt>\n"
                "       ref_driver model <name> <n>     (prints the text of a built-in model, simbody_b200/host/model_spec.h)\n"
                "       ref_driver eval|energy <model.txt> <in.bin> <out.bin> <N>\n"
                "       ref_driver step <model.txt> <in.bin> <out.bin> <N> <h> <nsteps> [accuracy] [consTol] [infNorm] [projectEveryStep]\n"
                "       ref_driver adaptive <model.txt> <in.bin> <out.bin> <N> <tFinal> [accuracy] [allowInterpolation] [consTol] [infNorm] [projectEveryStep]\n"
                "       ref_driver bench <model.txt> <in.bin> <N> <h> <nsteps> <threads> [out.bin]\n");
            return 2;
        }
        const std::string cmd = argv[1];
        if (cmd == "model") {       // the built-in model texts, so that a caller needs no other library to obtain them
            if (argc < 4) return 2;
            std::fputs(sbk::toText(sbk::makeNamedModel(argv[2], std::atoi(argv[3]))).c_str(), stdout);
            return 0;
        }
        const sbk::ModelSpec spec = sbk::fromText(slurp(argv[2]));
        if (cmd == "bench") {
            if (argc < 8) return 2;
            return cmdBench(spec, argv[3], std::atoi(argv[4]), std::atof(argv[5]), std::atoi(argv[6]),
                            std::atoi(argv[7]), argc > 8 ? argv[8] : nullptr);
        }
        RefSystem rs(spec);
        if (cmd == "lower") {
            sbk::ModelSpec low = sbk::lowerSimbodySystem(rs.system, rs.matter, &rs.forces, spec.name, &rs.defaultState);
            std::fputs(sbk::toText(low).c_str(), stdout);
            return 0;
        }
        if (cmd == "slots") {
            std::printf("nb %d nq %d nu %d nquat %d\n", rs.nb, rs.nq, rs.nu, rs.nquat);
            for (MobilizedBodyIndex b(0); b < rs.nb; ++b) {
                const MobilizedBody& m = rs.matter.getMobilizedBody(b);
                std::printf("body %d q0 %d nq %d u0 %d nu %d level %d\n", (int)b,
                    b == 0 ? 0 : (int)m.getFirstQIndex(rs.defaultState), m.getNumQ(rs.defaultState),
                    b == 0 ? 0 : (int)m.getFirstUIndex(rs.defaultState), m.getNumU(rs.defaultState),
                    m.getLevelInMultibodyTree());
            }
            return 0;
        }
        if (cmd == "eval" && argc >= 6) return cmdEval(rs, argv[3], argv[4], std::atoi(argv[5]));
        if (cmd == "energy" && argc >= 6) return cmdEnergy(rs, argv[3], argv[4], std::atoi(argv[5]));
        if (cmd == "extras" && argc >= 6) return cmdExtras(rs, argv[3], argv[4], std::atoi(argv[5]));
        auto optsFrom = [&](int first) { IntegOpts o; if (argc > first) o.consTol = std::atof(argv[first]);
                                         if (argc > first + 1) o.infNorm = std::atoi(argv[first + 1]); if (argc > first + 2) o.projectEvery = std::atoi(argv[first + 2]); return o; };
        if (cmd == "step" && argc >= 8)
            return cmdStep(rs, argv[3], argv[4], std::atoi(argv[5]), std::atof(argv[6]), std::atoi(argv[7]),
                           argc > 8 ? std::atof(argv[8]) : -1, optsFrom(9));
        if (cmd == "adaptive" && argc >= 7)
            return cmdAdaptive(rs, argv[3], argv[4], std::atoi(argv[5]), std::atof(argv[6]), argc > 7 ? std::atof(argv[7]) : -1,
                               argc > 8 ? std::atoi(argv[8]) != 0 : true, optsFrom(9));
        std::fprintf(stderr, "ref_driver: bad command line\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_driver: exception: %s\n", e.what());
        return 1;
    }
}
