// BatchedMatter.h -- C++ host facade over the C ABI (include/sbk.h).
//
// Mirrors the reference's operator surface for the hot path with a leading batch dimension:
//   SimbodyMatterSubsystem::realizePositionKinematics / realizeVelocityKinematics /
//   realizeArticulatedBodyInertias            (SimbodyMatterSubsystem.h:2691,2705,2730)
//   calcAcceleration / calcAccelerationIgnoringConstraints   (:2141,:2171)
//   multiplyByM / multiplyByMInv              (:1262,:1343)
//   calcResidualForceIgnoringConstraints      (:2234)
//   calcMobilizerReactionForces, multiplyBySystemJacobian[Transpose]   (:2479,:554,:646)
//   RungeKuttaMersonIntegrator + Integrator::setFixedStepSize / setAccuracy /
//   setConstraintTolerance / setUseInfinityNorm / setProjectEveryStep / stepBy / getNumStepsTaken /
//   getNumRealizations                        (simmath/Integrator.h:143-394)
// Error behaviour follows the reference: wrong vector lengths and stage violations throw
// (std::invalid_argument / std::logic_error) instead of SimTK exceptions.
// Batched vectors are slot-major: element (slot i, instance k) at v[i*N + k].
// Header-only; link with libsbk.so.  No Simbody headers are needed here (the lowering that does
// need them is lower_simbody.h).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "sbk.h"
#include "model_spec.h"

namespace sbk {

inline void throwOnError(int rc) {
    if (rc == SBK_OK) return;
    const std::string msg = sbk_last_error();
    if (rc == SBK_ERR_ARG)   throw std::invalid_argument(msg);
    if (rc == SBK_ERR_STAGE) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

class Topology {
public:
    explicit Topology(const ModelSpec& spec)
    :   h_(sbk_topology_create_ex(spec.bodies.data(), (int)spec.bodies.size(),
                                  spec.forces.empty() ? nullptr : spec.forces.data(), (int)spec.forces.size(),
                                  spec.useEulerAngles ? SBK_TOPOLOGY_EULER_ANGLES : 0u)) {
        if (!h_) throw std::runtime_error(sbk_last_error());
        throwOnError(sbk_topology_counts(h_, &nb_, &nq_, &nu_, &nquat_, &nlevels_));
    }
    ~Topology() { sbk_topology_destroy(h_); }
    Topology(const Topology&) = delete; Topology& operator=(const Topology&) = delete;
    int getNumBodies() const { return nb_; }  int getNQ() const { return nq_; }  int getNU() const { return nu_; }
    int getNumQuaternions() const { return nquat_; }  int getNumLevels() const { return nlevels_; }
    const sbk_topology* handle() const { return h_; }
private:
    sbk_topology* h_; int nb_ = 0, nq_ = 0, nu_ = 0, nquat_ = 0, nlevels_ = 0;
};

// N independent instances of one lowered system, resident on one GPU.
class BatchedMatter {
public:
    BatchedMatter(const Topology& topo, int nInstances, int device = 0, void* cudaStream = nullptr)
    :   topo_(topo), n_(nInstances), h_(sbk_batch_create(topo.handle(), nInstances, device, cudaStream)) {
        if (!h_) throw std::runtime_error(sbk_last_error());
    }
    ~BatchedMatter() { sbk_batch_destroy(h_); }
    BatchedMatter(const BatchedMatter&) = delete; BatchedMatter& operator=(const BatchedMatter&) = delete;

    int getNumInstances() const { return n_; }
    const Topology& getTopology() const { return topo_; }
    sbk_batch* handle() { return h_; }

    // State::updQ / updU (State.h:962-1043)
    void setState(const std::vector<double>& q, const std::vector<double>& u) {
        need(q, topo_.getNQ(), "q"); need(u, topo_.getNU(), "u");
        throwOnError(sbk_set_state(h_, q.data(), u.data(), nullptr));
    }
    void getState(std::vector<double>& q, std::vector<double>& u) {
        q.resize((size_t)topo_.getNQ()*n_); u.resize((size_t)topo_.getNU()*n_);
        throwOnError(sbk_get_state(h_, q.data(), u.data(), nullptr));
    }
    void realizePositionKinematics()      { throwOnError(sbk_realize_position(h_)); }
    void realizeVelocityKinematics()      { throwOnError(sbk_realize_velocity(h_)); }
    void realizeArticulatedBodyInertias() { throwOnError(sbk_realize_articulated_body_inertias(h_)); }
    void realizeAcceleration()            { throwOnError(sbk_realize_acceleration(h_)); }   // System::realize(s, Stage::Acceleration)
    void getUDot(std::vector<double>& udot) { udot.resize((size_t)topo_.getNU()*n_); throwOnError(sbk_get_udot(h_, udot.data())); }
    void getQDot(std::vector<double>& qdot) { qdot.resize((size_t)topo_.getNQ()*n_); throwOnError(sbk_get_qdot(h_, qdot.data())); }
    void getQDotDot(std::vector<double>& qdd) { qdd.resize((size_t)topo_.getNQ()*n_); throwOnError(sbk_get_qdotdot(h_, qdd.data())); }
    // MobilizedBody::getBodyTransform / getBodyVelocity / getBodyAcceleration for every body (MobilizedBody.h:560-620):
    // X_GB [nb][12][N] (R row-major, p), V_GB / A_GB [nb][6][N] (angular, linear), expressed in Ground
    void getBodyTransforms(std::vector<double>& X_GB)    { X_GB.resize((size_t)12*topo_.getNumBodies()*n_); throwOnError(sbk_get_body_transforms(h_, X_GB.data())); }
    void getBodyVelocities(std::vector<double>& V_GB)    { V_GB.resize((size_t)6*topo_.getNumBodies()*n_); throwOnError(sbk_get_body_velocities(h_, V_GB.data())); }
    void getBodyAccelerations(std::vector<double>& A_GB) { A_GB.resize((size_t)6*topo_.getNumBodies()*n_); throwOnError(sbk_get_body_accelerations(h_, A_GB.data())); }
    // MultibodySystem::calcKineticEnergy / calcPotentialEnergy (MultibodySystem.h:143-168), one value per instance
    void calcEnergy(std::vector<double>& kinetic, std::vector<double>& potential) {
        kinetic.resize((size_t)n_); potential.resize((size_t)n_); throwOnError(sbk_calc_energy(h_, kinetic.data(), potential.data()));
    }
    // per-instance status words (include/sbk.h: bit 0 non-finite, bit 1 singular D, bit 2 attempt budget); returns the number of bad instances
    long long getStatus(std::vector<int32_t>& status) { status.resize((size_t)n_); int64_t bad = 0; throwOnError(sbk_get_status(h_, status.data(), &bad)); return bad; }

    // Zero-length force vectors mean "all zero", as in the reference (SimbodyMatterSubsystemRep.cpp:5546-5564).
    void calcAcceleration(const std::vector<double>& appliedMobilityForces, const std::vector<double>& appliedBodyForces,
                          std::vector<double>& udot, std::vector<double>& A_GB) {
        opt(appliedMobilityForces, topo_.getNU(), "appliedMobilityForces"); opt(appliedBodyForces, 6*topo_.getNumBodies(), "appliedBodyForces");
        udot.resize((size_t)topo_.getNU()*n_); A_GB.resize((size_t)6*topo_.getNumBodies()*n_);
        throwOnError(sbk_calc_acceleration(h_, ptr(appliedMobilityForces), ptr(appliedBodyForces), udot.data(), A_GB.data()));
    }
    void calcAccelerationIgnoringConstraints(const std::vector<double>& f, const std::vector<double>& F,
                                             std::vector<double>& udot, std::vector<double>& A_GB) { calcAcceleration(f, F, udot, A_GB); }
    void multiplyByM(const std::vector<double>& a, std::vector<double>& Ma) {
        need(a, topo_.getNU(), "a"); Ma.resize(a.size()); throwOnError(sbk_multiply_by_M(h_, a.data(), Ma.data()));
    }
    void multiplyByMInv(const std::vector<double>& v, std::vector<double>& MinvV) {
        need(v, topo_.getNU(), "v"); MinvV.resize(v.size()); throwOnError(sbk_multiply_by_MInv(h_, v.data(), MinvV.data()));
    }
    // SimbodyMatterSubsystem::calcMobilizerReactionForces (:2479): FM_G [nb][6][N], needs realizeAcceleration()
    void calcMobilizerReactionForces(std::vector<double>& FM_G) {
        FM_G.resize((size_t)6*topo_.getNumBodies()*n_); throwOnError(sbk_calc_mobilizer_reaction_forces(h_, FM_G.data()));
    }
    // calcCompositeBodyInertias: R [nb][10][N] = mass, com(3), unit inertia xx yy zz xy xz yz
    void calcCompositeBodyInertias(std::vector<double>& R) {
        R.resize((size_t)10*topo_.getNumBodies()*n_); throwOnError(sbk_calc_composite_body_inertias(h_, R.data()));
    }
    // multiplyBySystemJacobian / multiplyBySystemJacobianTranspose (:554,:646)
    void multiplyBySystemJacobian(const std::vector<double>& v, std::vector<double>& Jv) {
        need(v, topo_.getNU(), "v"); Jv.resize((size_t)6*topo_.getNumBodies()*n_);
        throwOnError(sbk_multiply_by_system_jacobian(h_, v.data(), Jv.data()));
    }
    void multiplyBySystemJacobianTranspose(const std::vector<double>& F_G, std::vector<double>& JtF) {
        need(F_G, 6*topo_.getNumBodies(), "F_G"); JtF.resize((size_t)topo_.getNU()*n_);
        throwOnError(sbk_multiply_by_system_jacobian_transpose(h_, F_G.data(), JtF.data()));
    }
    void calcResidualForceIgnoringConstraints(const std::vector<double>& appliedMobilityForces, const std::vector<double>& appliedBodyForces,
                                              const std::vector<double>& knownUdot, std::vector<double>& residual) {
        opt(appliedMobilityForces, topo_.getNU(), "appliedMobilityForces"); opt(appliedBodyForces, 6*topo_.getNumBodies(), "appliedBodyForces");
        opt(knownUdot, topo_.getNU(), "knownUdot");
        residual.resize((size_t)topo_.getNU()*n_);
        throwOnError(sbk_calc_residual_force(h_, ptr(appliedMobilityForces), ptr(appliedBodyForces), ptr(knownUdot), residual.data()));
    }
private:
    static const double* ptr(const std::vector<double>& v) { return v.empty() ? nullptr : v.data(); }
    void need(const std::vector<double>& v, int rows, const char* name) const {
        if (v.size() != (size_t)rows*n_) throw std::invalid_argument(std::string(name) + " has the wrong length");   // SimTK_APIARGCHECK
    }
    void opt(const std::vector<double>& v, int rows, const char* name) const { if (!v.empty()) need(v, rows, name); }
    const Topology& topo_; int n_; sbk_batch* h_;
};

// RungeKuttaMersonIntegrator for the whole batch: fixed steps (setFixedStepSize + stepBy) or error-controlled stepping to a
// time (stepTo; every instance keeps its own step-size history, Integrator.h:226).
class BatchedRungeKuttaMerson {
public:
    explicit BatchedRungeKuttaMerson(BatchedMatter& m) : m_(m) { sbk_rkm_default_opts(&o_); sbk_adaptive_default_opts(&ao_); }
    void setFixedStepSize(double h)        { h_ = h; }
    void setAccuracy(double acc)           { o_.accuracy = ao_.accuracy = acc; if (!userTol_) o_.constraint_tol = ao_.constraint_tol = acc/10; }   // IntegratorRep.h:737-740
    void setConstraintTolerance(double t)  { o_.constraint_tol = ao_.constraint_tol = t; userTol_ = true; }
    void setUseInfinityNorm(bool b)        { o_.use_infinity_norm = ao_.use_infinity_norm = b; }
    void setProjectEveryStep(bool b)       { o_.project_every_step = ao_.project_every_step = b; }
    void setInitialStepSize(double h)      { ao_.init_step = h; }                 // Integrator.h:322
    void setMinimumStepSize(double h)      { ao_.min_step = h; }                  // :327
    void setMaximumStepSize(double h)      { ao_.max_step = h; }                  // :332
    void setAllowInterpolation(bool b)     { ao_.allow_interpolation = b; }       // :357 (default here false = TimeStepper semantics)
    // Integrator::stepTo(reportTime) with error control (Integrator.h:226): every instance advances to tFinal with its own steps
    void stepTo(double tFinal) {
        const size_t n = (size_t)m_.getNumInstances();
        steps_.resize(n); attempts_.resize(n); lastStep_.resize(n);
        throwOnError(sbk_rkm_adaptive(m_.handle(), tFinal, &ao_, steps_.data(), attempts_.data(), lastStep_.data()));
    }
    // per-instance counters of the error-controlled run since the last setState (Integrator.h:286-290, :262)
    const std::vector<int32_t>& getNumStepsTakenPerInstance() const     { return steps_; }
    const std::vector<int32_t>& getNumStepsAttemptedPerInstance() const { return attempts_; }
    const std::vector<double>&  getPreviousStepSizeTaken() const        { return lastStep_; }
    long long getNumStepsAttempted() const { long long s = 0; for (int32_t a : attempts_) s += a; return s; }
    std::vector<double> getTime() { std::vector<double> t((size_t)m_.getNumInstances()); throwOnError(sbk_get_state(m_.handle(), nullptr, nullptr, t.data())); return t; }
    // nsteps accepted steps of size h for every instance; the state stays on the GPU.
    void stepBy(int nsteps) {
        if (!(h_ > 0)) throw std::logic_error("BatchedRungeKuttaMerson: call setFixedStepSize() first");
        throwOnError(sbk_rkm_step(m_.handle(), h_, nsteps, &o_, nullptr));
    }
    std::vector<double> getLastErrorNorms() { std::vector<double> e((size_t)m_.getNumInstances()); throwOnError(sbk_rkm_step(m_.handle(), h_ > 0 ? h_ : 1.0, 0, &o_, e.data())); return e; }
    long long getNumStepsTaken()    { int64_t s, r, p; throwOnError(sbk_rkm_stats(m_.handle(), &s, &r, &p)); return s; }
    long long getNumRealizations()  { int64_t s, r, p; throwOnError(sbk_rkm_stats(m_.handle(), &s, &r, &p)); return r; }
    long long getNumQProjections()  { int64_t s, r, p; throwOnError(sbk_rkm_stats(m_.handle(), &s, &r, &p)); return p; }
private:
    BatchedMatter& m_; sbk_rkm_opts o_; sbk_adaptive_opts ao_; double h_ = -1; bool userTol_ = false;
    std::vector<int32_t> steps_, attempts_; std::vector<double> lastStep_;
};

} // namespace sbk
