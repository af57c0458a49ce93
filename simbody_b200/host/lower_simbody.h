// lower_simbody.h -- lower a realized SimTK::MultibodySystem into a sbk::ModelSpec
// (the flat body table + force list that sbk_topology_create consumes).
//
// This is the piece a Simbody user adds to switch the hot path to the B200 engine: build
// the system exactly as before, call system.realizeTopology(), then
//     sbk::ModelSpec spec = sbk::lowerSimbodySystem(system, matter, forces);
//     sbk_topology* topo = sbk_topology_create(spec.bodies.data(), spec.bodies.size(), ...);
// Only the PUBLIC Simbody API is used:
//   SimbodyMatterSubsystem::getNumBodies / getMobilizedBody
//   MobilizedBody::getParentMobilizedBody (MobilizedBody.h:1627), getDefaultInboardFrame /
//   getDefaultOutboardFrame (:1605,:1609), getDefaultMassProperties (:1556),
//   MobilizedBody::Pin::isInstanceOf etc. (SimTKcommon PrivateImplementation.h:348),
//   Force::Gravity::getDefaultDownDirection / getDefaultMagnitude (Force_Gravity.h:276-278),
//   Force::MobilityLinearSpring::getDefaultStiffness / getDefaultQZero
//   (Force_MobilityLinearSpring.h:109,113), Force::MobilityLinearDamper::getDefaultDamping
//   (Force_MobilityLinearDamper.h:87), Force::UniformGravity::getGravity / getZeroHeight (Force.h:385-387);
//   Force::GlobalDamper has no getter: its constant is probed through Force::calcForceContribution.
// The spring/damper classes do not expose WHICH mobility they act on, so it is recovered by
// probing Force::calcForceContribution (Force.h:130) at a state where every q-q0 (resp. u)
// is non-zero: exactly one mobility force entry is non-zero.
//
// Requires the Simbody headers; it is compiled only into programs that link Simbody
// (oracle/ref_driver.cpp here, the user's program in production).  Unsupported features
// (other mobilizers, reversed mobilizers, constraints, other force types)
// raise std::runtime_error rather than being silently dropped.
#pragma once
#include <string>
#include "Simbody.h"
#include "simbody_b200/host/model_spec.h"

namespace sbk {

inline void lowerTransform(const SimTK::Transform& X, double out[12]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[3*i+j] = X.R()[i][j];
    for (int i = 0; i < 3; ++i) out[9+i] = X.p()[i];
}

inline ModelSpec lowerSimbodySystem(const SimTK::MultibodySystem&        system,
                                    const SimTK::SimbodyMatterSubsystem& matter,
                                    const SimTK::GeneralForceSubsystem*  forces,
                                    const std::string& name = "lowered",
                                    const SimTK::State* modelState = nullptr)   // where the modeling options live (Euler-angle mode)
{
    using namespace SimTK;
    if (!system.systemTopologyHasBeenRealized())
        throw std::runtime_error("lowerSimbodySystem: call system.realizeTopology() first");
    if (matter.getNumConstraints() != 0)
        throw std::runtime_error("lowerSimbodySystem: constraints are out of scope (tree topologies only)");
    if (matter.getNumParticles() != 0)
        throw std::runtime_error("lowerSimbodySystem: particles are not supported");

    ModelSpec spec; spec.name = name;
    const int nb = matter.getNumBodies();
    State state = system.getDefaultState();
    spec.useEulerAngles = modelState ? matter.getUseEulerAngles(*modelState) : matter.getUseEulerAngles(state);

    std::vector<int> uFirst(nb, 0), uCount(nb, 0);
    for (MobilizedBodyIndex mbx(0); mbx < nb; ++mbx) {
        const MobilizedBody& mobod = matter.getMobilizedBody(mbx);
        sbk_body_desc b = groundBody();
        if (mbx == 0) { spec.bodies.push_back(b); continue; }
        b.parent = (int)mobod.getParentMobilizedBody().getMobilizedBodyIndex();
        if      (MobilizedBody::Pin::isInstanceOf(mobod))       b.joint_type = SBK_JOINT_PIN;
        else if (MobilizedBody::Slider::isInstanceOf(mobod))    b.joint_type = SBK_JOINT_SLIDER;
        else if (MobilizedBody::Universal::isInstanceOf(mobod)) b.joint_type = SBK_JOINT_UNIVERSAL;
        else if (MobilizedBody::Ball::isInstanceOf(mobod))      b.joint_type = SBK_JOINT_BALL;
        else if (MobilizedBody::Free::isInstanceOf(mobod))      b.joint_type = SBK_JOINT_FREE;
        else if (MobilizedBody::Weld::isInstanceOf(mobod))      b.joint_type = SBK_JOINT_WELD;
        else if (MobilizedBody::Translation::isInstanceOf(mobod)) b.joint_type = SBK_JOINT_TRANSLATION;
        else if (MobilizedBody::Cylinder::isInstanceOf(mobod))  b.joint_type = SBK_JOINT_CYLINDER;
        else if (MobilizedBody::Planar::isInstanceOf(mobod))    b.joint_type = SBK_JOINT_PLANAR;
        else if (MobilizedBody::Gimbal::isInstanceOf(mobod))    b.joint_type = SBK_JOINT_GIMBAL;
        else throw std::runtime_error("lowerSimbodySystem: body " + std::to_string((int)mbx) +
                                      " uses a mobilizer outside {Pin,Slider,Universal,Ball,Free,Weld,Translation,Cylinder,Planar,Gimbal}");
        // NOTE: the public API has no getter for MobilizedBody::Direction; reversed mobilizers
        // are out of scope and must not be used with this lowering.
        const MassProperties& mp = mobod.getDefaultMassProperties();
        b.mass = mp.getMass();
        for (int k = 0; k < 3; ++k) b.com_B[k] = mp.getMassCenter()[k];
        const Vec3& mom = mp.getUnitInertia().getMoments();
        const Vec3& prd = mp.getUnitInertia().getProducts();   // xy, xz, yz
        for (int k = 0; k < 3; ++k) { b.unit_inertia_OB_B[k] = mom[k]; b.unit_inertia_OB_B[3+k] = prd[k]; }
        lowerTransform(mobod.getDefaultInboardFrame(),  b.X_PF);
        lowerTransform(mobod.getDefaultOutboardFrame(), b.X_BM);
        uFirst[mbx] = (int)mobod.getFirstUIndex(state); uCount[mbx] = mobod.getNumU(state);
        if (mobod.getNumQ(state) != jointNQ(b.joint_type) || uCount[mbx] != jointNU(b.joint_type))
            throw std::runtime_error("lowerSimbodySystem: unexpected nq/nu for body " + std::to_string((int)mbx));
        spec.bodies.push_back(b);
    }

    if (forces) {
        auto slotToBodyCoord = [&](int uslot, int& body, int& coord) {
            for (int b = 1; b < nb; ++b)
                if (uslot >= uFirst[b] && uslot < uFirst[b] + uCount[b]) { body = b; coord = uslot - uFirst[b]; return; }
            throw std::runtime_error("lowerSimbodySystem: mobility slot not found");
        };
        for (ForceIndex fx(0); fx < forces->getNumForces(); ++fx) {
            const Force& f = forces->getForce(fx);
            if (forces->isForceDisabled(state, fx)) continue;
            if (Force::Gravity::isInstanceOf(f)) {
                const Force::Gravity& g = Force::Gravity::downcast(f);
                for (MobilizedBodyIndex mbx(1); mbx < nb; ++mbx)
                    if (g.getDefaultBodyIsExcluded(mbx))
                        throw std::runtime_error("lowerSimbodySystem: per-body gravity exclusion is not supported");
                const UnitVec3& d = g.getDefaultDownDirection();
                spec.forces.push_back(gravityForce(g.getDefaultMagnitude(), d[0], d[1], d[2]));
            } else if (Force::MobilityLinearSpring::isInstanceOf(f)) {
                const Force::MobilityLinearSpring& s = Force::MobilityLinearSpring::downcast(f);
                const Real k = s.getDefaultStiffness(), q0 = s.getDefaultQZero();
                State probe = system.getDefaultState();
                probe.updQ() = q0 + 1; probe.updU() = 0;
                system.realize(probe, Stage::Velocity);
                Vector_<SpatialVec> bf; Vector_<Vec3> pf; Vector mf;
                f.calcForceContribution(probe, bf, pf, mf);
                int hit = -1, nhit = 0;
                for (int i = 0; i < mf.size(); ++i) if (mf[i] != 0) { hit = i; ++nhit; }
                if (nhit != 1) throw std::runtime_error("lowerSimbodySystem: could not locate spring mobility");
                int body, coord; slotToBodyCoord(hit, body, coord);
                const int jt = spec.bodies[body].joint_type;
                if (jt == SBK_JOINT_BALL || jt == SBK_JOINT_FREE)
                    throw std::runtime_error("lowerSimbodySystem: MobilityLinearSpring on a quaternion mobilizer "
                                             "is not supported (reference mixes q and u indices, Force.cpp:348)");
                spec.forces.push_back(springForce(body, coord, k, q0));
            } else if (Force::MobilityLinearDamper::isInstanceOf(f)) {
                const Force::MobilityLinearDamper& d = Force::MobilityLinearDamper::downcast(f);
                State probe = system.getDefaultState();
                probe.updU() = 1;
                system.realize(probe, Stage::Velocity);
                Vector_<SpatialVec> bf; Vector_<Vec3> pf; Vector mf;
                f.calcForceContribution(probe, bf, pf, mf);
                int hit = -1, nhit = 0;
                for (int i = 0; i < mf.size(); ++i) if (mf[i] != 0) { hit = i; ++nhit; }
                if (nhit != 1) throw std::runtime_error("lowerSimbodySystem: could not locate damper mobility");
                int body, coord; slotToBodyCoord(hit, body, coord);
                spec.forces.push_back(damperForce(body, coord, d.getDefaultDamping()));
            } else if (Force::MobilityConstantForce::isInstanceOf(f)) {
                const Force::MobilityConstantForce& mc = Force::MobilityConstantForce::downcast(f);
                const Real f0 = mc.getDefaultForce();
                if (f0 == 0) continue;                              // contributes nothing; its mobility cannot be located
                State probe = system.getDefaultState();
                system.realize(probe, Stage::Velocity);
                Vector_<SpatialVec> bf; Vector_<Vec3> pf; Vector mf;
                f.calcForceContribution(probe, bf, pf, mf);
                int hit = -1, nhit = 0;
                for (int i = 0; i < mf.size(); ++i) if (mf[i] != 0) { hit = i; ++nhit; }
                if (nhit != 1) throw std::runtime_error("lowerSimbodySystem: could not locate MobilityConstantForce mobility");
                int body, coord; slotToBodyCoord(hit, body, coord);
                spec.forces.push_back(mobilityConstantForce(body, coord, f0));
            } else if (Force::UniformGravity::isInstanceOf(f)) {
                const Force::UniformGravity& g = Force::UniformGravity::downcast(f);
                if (g.getZeroHeight() != 0)
                    throw std::runtime_error("lowerSimbodySystem: UniformGravity with a non-zero zero-height is not supported");
                const Vec3 gv = g.getGravity();
                spec.forces.push_back(uniformGravityForce(gv[0], gv[1], gv[2]));
            } else if (Force::GlobalDamper::isInstanceOf(f)) {
                // no getter for the damping constant: probe it (f = -c*u at u = 1 on every mobility)
                State probe = system.getDefaultState();
                probe.updU() = 1;
                system.realize(probe, Stage::Velocity);
                Vector_<SpatialVec> bf; Vector_<Vec3> pf; Vector mf;
                f.calcForceContribution(probe, bf, pf, mf);
                if (mf.size() == 0) throw std::runtime_error("lowerSimbodySystem: GlobalDamper on a system without mobilities");
                for (int i = 1; i < mf.size(); ++i) if (mf[i] != mf[0]) throw std::runtime_error("lowerSimbodySystem: unexpected GlobalDamper response");
                spec.forces.push_back(globalDamperForce(-mf[0]));
            } else if (Force::TwoPointLinearSpring::isInstanceOf(f) || Force::TwoPointLinearDamper::isInstanceOf(f)) {
                // Neither class has a getter (Force.h:231-262): bodies, stations and constants are identified from
                // Force::calcForceContribution at a few seeded states.  The element loads body1 with (s1_G x f, f) and body2 with
                // -(s2_G x f, f), f along the line through the two stations -- so in each body's own frame every probe's line of
                // action passes through that body's station: a 3x3 least-squares intersection per body, then a linear fit of
                // the force magnitude against the stretch (spring) or the closing speed (damper).  Recovered values are snapped to 12
                // significant digits (user-entered parameters come back exactly).
                const bool isSpring = Force::TwoPointLinearSpring::isInstanceOf(f);
                auto snap = [](double x) { if (std::fabs(x) < 1e-12) return 0.0; char buf[40]; std::snprintf(buf, sizeof buf, "%.12g", x); return std::strtod(buf, nullptr); };
                XorShift64 rng(0x51ED27A1B0C3ull + (uint64_t)(int)fx);
                const int K = 8;
                int b1 = -1, b2 = -1;
                Mat33 A[2] = {Mat33(0), Mat33(0)}; Vec3 rhs[2] = {Vec3(0), Vec3(0)};
                std::vector<State> probes; std::vector<Vec3> f1s;
                for (int i = 0; i < K; ++i) {
                    State probe = system.getDefaultState();
                    for (int j = 0; j < probe.getNQ(); ++j) probe.updQ()[j] += 0.7*rng.next();
                    for (int j = 0; j < probe.getNU(); ++j) probe.updU()[j] = rng.next();
                    for (MobilizedBodyIndex mbx(1); mbx < nb; ++mbx) {          // unit quaternions
                        const int jt = spec.bodies[mbx].joint_type;
                        if ((jt == SBK_JOINT_BALL || jt == SBK_JOINT_FREE) && !spec.useEulerAngles) {
                            const int q0 = (int)matter.getMobilizedBody(mbx).getFirstQIndex(probe);
                            Real n2 = 0; for (int j = 0; j < 4; ++j) n2 += probe.getQ()[q0 + j]*probe.getQ()[q0 + j];
                            for (int j = 0; j < 4; ++j) probe.updQ()[q0 + j] /= std::sqrt(n2);
                        }
                    }
                    system.realize(probe, Stage::Velocity);
                    Vector_<SpatialVec> bf; Vector_<Vec3> pf; Vector mf;
                    f.calcForceContribution(probe, bf, pf, mf);
                    int hits[2], nh = 0;
                    for (int b = 0; b < bf.size(); ++b) if (bf[b][1].norm() > 0) { if (nh < 2) hits[nh] = b; ++nh; }
                    if (nh != 2) throw std::runtime_error("lowerSimbodySystem: a two-point force element must load exactly two bodies");
                    if (b1 < 0) { b1 = hits[0]; b2 = hits[1]; }
                    else if (b1 != hits[0] || b2 != hits[1]) throw std::runtime_error("lowerSimbodySystem: inconsistent two-point force element");
                    for (int w = 0; w < 2; ++w) {
                        const MobilizedBody& mb = matter.getMobilizedBody(MobilizedBodyIndex(w ? b2 : b1));
                        const Rotation& R = mb.getBodyTransform(probe).R();
                        const Vec3 fG = bf[w ? b2 : b1][1], mG = bf[w ? b2 : b1][0];
                        const Vec3 aB = ~R*((fG % mG)/fG.normSqr()), dB = ~R*(fG/fG.norm());
                        const Mat33 Pm = Mat33(1) - Mat33(dB*~dB);
                        A[w] += Pm; rhs[w] += Pm*aB;
                    }
                    probes.push_back(probe); f1s.push_back(bf[b1][1]);
                }
                Vec3 st[2];
                for (int w = 0; w < 2; ++w) { st[w] = A[w].invert()*rhs[w]; for (int j = 0; j < 3; ++j) st[w][j] = snap(st[w][j]); }
                // magnitude law: f1 = frc * unit(p2 - p1), frc = k (d - x0)  |  c (vRel . unit)
                Real sxx = 0, sx = 0, sy = 0, sxy = 0;
                for (int i = 0; i < K; ++i) {
                    system.realize(probes[i], Stage::Velocity);          // a copied State keeps its variables, not its realized cache
                    const MobilizedBody& m1 = matter.getMobilizedBody(MobilizedBodyIndex(b1)); const MobilizedBody& m2 = matter.getMobilizedBody(MobilizedBodyIndex(b2));
                    const Vec3 r = m2.findStationLocationInGround(probes[i], st[1]) - m1.findStationLocationInGround(probes[i], st[0]);
                    const Real d = r.norm(), frc = dot(f1s[i], r/d);
                    const Real x = isSpring ? d : dot(m2.findStationVelocityInGround(probes[i], st[1]) - m1.findStationVelocityInGround(probes[i], st[0]), r/d);
                    sxx += x*x; sx += x; sy += frc; sxy += x*frc;
                }
                if (isSpring) {
                    const Real k = (K*sxy - sx*sy)/(K*sxx - sx*sx), c0 = (sy - k*sx)/K;          // frc = k d + c0, x0 = -c0 / k
                    const double s1[3] = {st[0][0], st[0][1], st[0][2]}, s2[3] = {st[1][0], st[1][1], st[1][2]};
                    spec.forces.push_back(twoPointSpringForce(b1, s1, b2, s2, snap(k), snap(-c0/k)));
                } else {
                    const double s1[3] = {st[0][0], st[0][1], st[0][2]}, s2[3] = {st[1][0], st[1][1], st[1][2]};
                    spec.forces.push_back(twoPointDamperForce(b1, s1, b2, s2, snap(sxy/sxx)));
                }
            } else {
                throw std::runtime_error("lowerSimbodySystem: force element " + std::to_string((int)fx) +
                                         " is outside {Gravity, UniformGravity, MobilityLinearSpring, MobilityLinearDamper, MobilityConstantForce, GlobalDamper, "
                                         "TwoPointLinearSpring, TwoPointLinearDamper}");
            }
        }
    }
    return spec;
}

} // namespace sbk
