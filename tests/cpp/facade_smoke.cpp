// facade_smoke.cpp -- exercises the C++ host facade (simbody_b200/host/BatchedMatter.h) end to end:
// built-in model -> Topology -> BatchedMatter -> operators -> fixed-step RKM.  Needs a GPU.
// Checks the reference's own invariants (Simbody/tests/TestMassMatrix.cpp:670-787):
// M^-1 (M v) = v and inverse dynamics of the forward-dynamics accelerations reproduces the forces.
#include <cmath>
#include <cstdio>
#include "simbody_b200/host/BatchedMatter.h"

int main() {
    try {
        const int N = 512;
        sbk::Topology topo(sbk::makeNamedModel("humanoid30", 0));
        sbk::BatchedMatter matter(topo, N);
        const int nq = topo.getNQ(), nu = topo.getNU(), nb = topo.getNumBodies();
        std::vector<double> q((size_t)nq*N, 0.0), u((size_t)nu*N, 0.0);
        sbk::XorShift64 rng(42);
        for (auto& x : q) x = 0.3*rng.next();
        for (auto& x : u) x = rng.next();
        // unit quaternions: pelvis Free (slots 0-3) and every Ball
        sbk::ModelSpec spec = sbk::makeNamedModel("humanoid30", 0);
        int slot = 0;
        for (int b = 1; b < nb; ++b) {
            const int jt = spec.bodies[b].joint_type;
            if (jt == SBK_JOINT_BALL || jt == SBK_JOINT_FREE)
                for (int k = 0; k < N; ++k) {
                    double n2 = 0; for (int i = 0; i < 4; ++i) n2 += q[(size_t)(slot+i)*N + k]*q[(size_t)(slot+i)*N + k];
                    q[(size_t)slot*N + k] += 1.0; n2 = 0;
                    for (int i = 0; i < 4; ++i) n2 += q[(size_t)(slot+i)*N + k]*q[(size_t)(slot+i)*N + k];
                    for (int i = 0; i < 4; ++i) q[(size_t)(slot+i)*N + k] /= std::sqrt(n2);
                }
            slot += sbk::jointNQ(jt);
        }
        matter.setState(q, u);
        bool threw = false;
        try { std::vector<double> t; matter.multiplyByM(u, t); } catch (const std::logic_error&) { threw = true; }   // stage check
        if (!threw) { std::printf("FAIL: stage violation not reported\n"); return 1; }
        matter.realizeVelocityKinematics();
        std::vector<double> v((size_t)nu*N), Mv, back, f((size_t)nu*N), F((size_t)6*nb*N), udot, A, resid;
        for (auto& x : v) x = rng.next(); for (auto& x : f) x = rng.next(); for (auto& x : F) x = rng.next();
        matter.multiplyByM(v, Mv); matter.multiplyByMInv(Mv, back);
        double e1 = 0; for (size_t i = 0; i < v.size(); ++i) e1 = std::fmax(e1, std::fabs(back[i] - v[i]));
        matter.calcAcceleration(f, F, udot, A);
        matter.calcResidualForceIgnoringConstraints(f, F, udot, resid);
        double e2 = 0; for (double r : resid) e2 = std::fmax(e2, std::fabs(r));
        // reactions, Jacobian products, composite inertias: <J v, F> == <v, ~J F>; the reaction on Ground balances gravity at rest
        std::vector<double> FM, Jv, JtF, R;
        matter.realizeAcceleration(); matter.calcMobilizerReactionForces(FM);
        matter.multiplyBySystemJacobian(v, Jv); matter.multiplyBySystemJacobianTranspose(F, JtF); matter.calcCompositeBodyInertias(R);
        double lhs = 0, rhs = 0;
        for (int i = 0; i < 6*nb; ++i) lhs += Jv[(size_t)i*N]*F[(size_t)i*N];
        for (int i = 0; i < nu; ++i)   rhs += v[(size_t)i*N]*JtF[(size_t)i*N];
        if (!(std::fabs(lhs - rhs) < 1e-9*(1 + std::fabs(lhs))) || FM.size() != (size_t)6*nb*N || R.size() != (size_t)10*nb*N) {
            std::printf("FAIL: Jacobian adjoint %.3e vs %.3e\n", lhs, rhs); return 1; }
        sbk::BatchedRungeKuttaMerson integ(matter);
        const long long r0 = integ.getNumRealizations();      // the explicit realizeAcceleration() above is counted too
        integ.setFixedStepSize(1e-3); integ.stepBy(10);
        std::printf("facade_smoke: |Minv(M v)-v|=%.3e  |ID(FD)|=%.3e  steps=%lld realizations=%lld\n", e1, e2,
                    integ.getNumStepsTaken(), integ.getNumRealizations());
        if (!(e1 < 1e-9 && e2 < 1e-8 && integ.getNumStepsTaken() == 10LL*N && integ.getNumRealizations() - r0 == 50LL*N)) { std::printf("FAIL\n"); return 1; }
        // getters, energies, status, error-controlled stepTo (Integrator.h:226,286-290): every instance lands on the report time
        std::vector<double> X, V, ke, pe; std::vector<int32_t> st;
        matter.realizeVelocityKinematics(); matter.getBodyTransforms(X); matter.getBodyVelocities(V); matter.calcEnergy(ke, pe);
        if (X.size() != (size_t)12*nb*N || V.size() != (size_t)6*nb*N || !(ke[0] > 0) || matter.getStatus(st) != 0) { std::printf("FAIL: getters\n"); return 1; }
        integ.setAccuracy(1e-4); integ.stepTo(0.05);
        const std::vector<double> t = integ.getTime();
        long long att = integ.getNumStepsAttempted(), stp = 0; for (int32_t a : integ.getNumStepsTakenPerInstance()) stp += a;
        for (double tk : t) if (tk != 0.05) { std::printf("FAIL: stepTo ended at %.17g\n", tk); return 1; }
        if (!(stp >= N && att >= stp)) { std::printf("FAIL: adaptive counters\n"); return 1; }
        std::printf("facade_smoke: stepTo(0.05) steps=%lld attempts=%lld\n", stp, att);
        std::printf("OK\n");
        return 0;
    } catch (const std::exception& e) { std::printf("FAIL: %s\n", e.what()); return 1; }
}
