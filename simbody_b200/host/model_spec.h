// model_spec.h -- neutral description of a tree-topology multibody model.
//
// A ModelSpec is the *input description* shared by (a) the engine's topology compiler
// (csrc/topology.cpp, behind sbk_topology_create) and (b) the Simbody-side harness
// (oracle/ref_driver.cpp), which builds a real SimTK::MultibodySystem from it through the
// public Simbody API and lowers it back with lower_simbody.h.  It carries exactly what the
// reference's mobilized-body constructors take (Simbody/include/simbody/internal/
// MobilizedBody_Pin.h etc.): parent, mobilizer kind, MassProperties, X_PF, X_BM; plus the
// in-scope force elements (Force_Gravity.h:70, Force_MobilityLinearSpring.h:66,
// Force_MobilityLinearDamper.h:61, Force::UniformGravity and Force::GlobalDamper, Force.h:356-390).
//
// Header-only, no dependencies beyond the C++ standard library and include/sbk.h.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sbk.h"

namespace sbk {

struct ModelSpec {
    std::string                 name;
    std::vector<sbk_body_desc>  bodies;   // [0] = Ground, MobilizedBodyIndex order
    std::vector<sbk_force_desc> forces;
    bool                        useEulerAngles = false;   // SimbodyMatterSubsystem::setUseEulerAngles: Ball / Free use x-y-z angles
};

inline int jointNQ(int jt) {
    switch (jt) { case SBK_JOINT_PIN: case SBK_JOINT_SLIDER: return 1; case SBK_JOINT_UNIVERSAL: case SBK_JOINT_CYLINDER: return 2;
                  case SBK_JOINT_TRANSLATION: case SBK_JOINT_PLANAR: case SBK_JOINT_GIMBAL: return 3;
                  case SBK_JOINT_BALL: return 4; case SBK_JOINT_FREE: return 7; default: return 0; }
}
inline int jointNU(int jt) {
    switch (jt) { case SBK_JOINT_PIN: case SBK_JOINT_SLIDER: return 1; case SBK_JOINT_UNIVERSAL: case SBK_JOINT_CYLINDER: return 2;
                  case SBK_JOINT_BALL: case SBK_JOINT_TRANSLATION: case SBK_JOINT_PLANAR: case SBK_JOINT_GIMBAL: return 3;
                  case SBK_JOINT_FREE: return 6; default: return 0; }
}
inline const char* jointName(int jt) {
    switch (jt) { case SBK_JOINT_GROUND: return "GROUND"; case SBK_JOINT_PIN: return "PIN";
                  case SBK_JOINT_SLIDER: return "SLIDER"; case SBK_JOINT_UNIVERSAL: return "UNIVERSAL";
                  case SBK_JOINT_BALL: return "BALL"; case SBK_JOINT_FREE: return "FREE"; case SBK_JOINT_WELD: return "WELD";
                  case SBK_JOINT_TRANSLATION: return "TRANSLATION"; case SBK_JOINT_CYLINDER: return "CYLINDER"; case SBK_JOINT_PLANAR: return "PLANAR";
                  case SBK_JOINT_GIMBAL: return "GIMBAL";
                  default: return "?"; }
}
inline int jointFromName(const std::string& s) {
    if (s == "GROUND") return SBK_JOINT_GROUND; if (s == "PIN") return SBK_JOINT_PIN;
    if (s == "SLIDER") return SBK_JOINT_SLIDER; if (s == "UNIVERSAL") return SBK_JOINT_UNIVERSAL;
    if (s == "BALL") return SBK_JOINT_BALL; if (s == "FREE") return SBK_JOINT_FREE;
    if (s == "WELD") return SBK_JOINT_WELD; if (s == "TRANSLATION") return SBK_JOINT_TRANSLATION;
    if (s == "CYLINDER") return SBK_JOINT_CYLINDER; if (s == "PLANAR") return SBK_JOINT_PLANAR;
    if (s == "GIMBAL") return SBK_JOINT_GIMBAL;
    throw std::runtime_error("unknown joint name " + s);
}

// ---- small helpers for building frames ----------------------------------------------------
inline void setIdentityX(double X[12]) {
    for (int i = 0; i < 12; ++i) X[i] = 0; X[0] = X[4] = X[8] = 1;
}
// Rodrigues rotation about (unnormalised) axis; row-major R into X[0..8], translation p.
inline void setAxisAngleX(double X[12], double angle, double ax, double ay, double az,
                          double px, double py, double pz) {
    const double n = std::sqrt(ax*ax + ay*ay + az*az);
    const double x = ax/n, y = ay/n, z = az/n, c = std::cos(angle), s = std::sin(angle), t = 1 - c;
    X[0] = t*x*x + c;   X[1] = t*x*y - s*z; X[2] = t*x*z + s*y;
    X[3] = t*x*y + s*z; X[4] = t*y*y + c;   X[5] = t*y*z - s*x;
    X[6] = t*x*z - s*y; X[7] = t*y*z + s*x; X[8] = t*z*z + c;
    X[9] = px; X[10] = py; X[11] = pz;
}
struct XorShift64 {   // the generator SURVEY.md Appendix C uses for the golden fixture
    uint64_t z;
    explicit XorShift64(uint64_t seed = 88172645463325252ull) : z(seed) {}
    double next() { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return double(z >> 11) / 9007199254740992.0 * 2 - 1; }
};

inline sbk_body_desc groundBody() {
    sbk_body_desc g; std::memset(&g, 0, sizeof g);
    g.parent = -1; g.joint_type = SBK_JOINT_GROUND; setIdentityX(g.X_PF); setIdentityX(g.X_BM);
    return g;
}
// Unit inertia about the body origin from a central inertia of a solid box (dims a,b,c) rotated
// by a small rotation and shifted to the origin by com.  Always physically valid.
inline void boxUnitInertia(double ui[6], double a, double b, double c, const double com[3],
                           double rotAngle, double rx, double ry, double rz) {
    const double Ic[3] = { (b*b + c*c)/12, (a*a + c*c)/12, (a*a + b*b)/12 };
    double X[12]; setAxisAngleX(X, rotAngle, rx, ry, rz, 0, 0, 0);
    // G = R diag(Ic) R^T
    double G[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        G[i][j] = 0; for (int k = 0; k < 3; ++k) G[i][j] += X[3*i+k]*Ic[k]*X[3*j+k];
    }
    const double c2 = com[0]*com[0] + com[1]*com[1] + com[2]*com[2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        G[i][j] += (i == j ? c2 : 0) - com[i]*com[j];
    ui[0] = G[0][0]; ui[1] = G[1][1]; ui[2] = G[2][2];
    ui[3] = 0.5*(G[0][1]+G[1][0]); ui[4] = 0.5*(G[0][2]+G[2][0]); ui[5] = 0.5*(G[1][2]+G[2][1]);
}
inline sbk_force_desc gravityForce(double g, double dx, double dy, double dz) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_GRAVITY; f.body = -1; f.a = g; f.dir[0] = dx; f.dir[1] = dy; f.dir[2] = dz; return f;
}
inline sbk_force_desc springForce(int body, int coord, double k, double q0) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_SPRING; f.body = body; f.coord = coord; f.a = k; f.b = q0; return f;
}
inline sbk_force_desc damperForce(int body, int coord, double c) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_DAMPER; f.body = body; f.coord = coord; f.a = c; return f;
}

inline sbk_force_desc uniformGravityForce(double gx, double gy, double gz) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_UNIFORM_GRAVITY; f.body = -1; f.a = 1; f.dir[0] = gx; f.dir[1] = gy; f.dir[2] = gz; return f;
}
inline sbk_force_desc mobilityConstantForce(int body, int coord, double f0) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_MOBILITY_CONSTANT; f.body = body; f.coord = coord; f.a = f0; return f;
}
inline sbk_force_desc twoPointSpringForce(int body1, const double s1[3], int body2, const double s2[3], double k, double x0) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_TWO_POINT_SPRING; f.body = body1; f.coord = body2; f.a = k; f.b = x0;
    for (int i = 0; i < 3; ++i) { f.dir[i] = s1[i]; f.station2[i] = s2[i]; }
    return f;
}
inline sbk_force_desc twoPointDamperForce(int body1, const double s1[3], int body2, const double s2[3], double c) {
    sbk_force_desc f = twoPointSpringForce(body1, s1, body2, s2, c, 0); f.kind = SBK_FORCE_TWO_POINT_DAMPER; return f;
}
inline sbk_force_desc globalDamperForce(double c) {
    sbk_force_desc f; std::memset(&f, 0, sizeof f);
    f.kind = SBK_FORCE_GLOBAL_DAMPER; f.body = -1; f.a = c; return f;
}

// ---- built-in models (BASELINE.json configs) -------------------------

// C1/C2/C3: README double pendulum generalised to n links (reference README.md:36-55):
// gravity -Y 9.8; every link m=1, com=0, UnitInertia(1); Pin with X_PF=I, X_BM=(I,(0,1,0)).
inline ModelSpec makePinChain(int n, const char* name = "pin_chain") {
    ModelSpec m; m.name = name; m.bodies.push_back(groundBody());
    for (int i = 1; i <= n; ++i) {
        sbk_body_desc b = groundBody();
        b.parent = i - 1; b.joint_type = SBK_JOINT_PIN; b.mass = 1.0;
        b.unit_inertia_OB_B[0] = b.unit_inertia_OB_B[1] = b.unit_inertia_OB_B[2] = 1.0;
        b.X_BM[10] = 1.0;   // p_BM = (0,1,0)
        m.bodies.push_back(b);
    }
    m.forces.push_back(gravityForce(9.8, 0, -1, 0));
    return m;
}

// SURVEY.md Appendix C mixed-joint fixture: Free->Ball->Universal->Pin->Slider + Pin branch.
inline ModelSpec makeMixed7() {
    ModelSpec m; m.name = "mixed7"; m.bodies.push_back(groundBody());
    const double mass = 2.3, com[3] = {0.01, -0.02, 0.03};
    // Inertia(3,4,5,.01,-.02,.04).shiftFromMassCenter(com, m), then / m  (TestMassMatrix.cpp:458-462)
    double I[6] = {3, 4, 5, 0.01, -0.02, 0.04};
    const double c2 = com[0]*com[0] + com[1]*com[1] + com[2]*com[2];
    I[0] += mass*(c2 - com[0]*com[0]); I[1] += mass*(c2 - com[1]*com[1]); I[2] += mass*(c2 - com[2]*com[2]);
    I[3] -= mass*com[0]*com[1]; I[4] -= mass*com[0]*com[2]; I[5] -= mass*com[1]*com[2];
    double XPF[12], XBM[12];
    setAxisAngleX(XPF,  0.3,  1, 2, 3,    0.1, -0.2, 0.3);
    setAxisAngleX(XBM, -0.4, -1, 0.5, 2,  0, 0.25, -0.1);
    const int jt[5] = {SBK_JOINT_FREE, SBK_JOINT_BALL, SBK_JOINT_UNIVERSAL, SBK_JOINT_PIN, SBK_JOINT_SLIDER};
    auto mk = [&](int parent, int joint) {
        sbk_body_desc b = groundBody(); b.parent = parent; b.joint_type = joint; b.mass = mass;
        for (int k = 0; k < 3; ++k) b.com_B[k] = com[k];
        for (int k = 0; k < 6; ++k) b.unit_inertia_OB_B[k] = I[k]*(1/mass);
        return b; };
    for (int i = 0; i < 5; ++i) {
        sbk_body_desc b = mk(i, jt[i]);
        std::memcpy(b.X_PF, XPF, sizeof XPF); std::memcpy(b.X_BM, XBM, sizeof XBM);
        m.bodies.push_back(b);
    }
    sbk_body_desc b6 = mk(1, SBK_JOINT_PIN);
    b6.X_PF[10] = -1.0; b6.X_BM[10] = 1.0;
    m.bodies.push_back(b6);
    m.forces.push_back(gravityForce(9.81, 0, -1, 0));
    m.forces.push_back(springForce(4, 0, 30.0, 0.2));
    m.forces.push_back(damperForce(5, 0, 1.5));
    return m;
}

// C4: 30-body humanoid.  Pelvis Free; spine torso Ball, chest Universal,
// neck Ball, head Pin, jaw Pin; per arm (off the chest) clavicle Pin, shoulder Ball, elbow Pin,
// forearm-twist Pin, wrist Universal, fingers Pin; per leg (off the pelvis) hip Ball, patella
// Slider, knee Pin, ankle Universal, subtalar Pin, toes Pin.  General (rotated) X_PF / X_BM on
// every mobilizer, anthropometric masses, full inertias; Gravity; spring+damper on every
// Pin/Slider/Universal coordinate; dampers on Ball speeds.
inline ModelSpec makeHumanoid30() {
    ModelSpec m; m.name = "humanoid30"; m.bodies.push_back(groundBody());
    XorShift64 rng(0x9E3779B97F4A7C15ull);
    struct Seg { int parent; int jt; double mass; double len; double wid; };
    std::vector<Seg> segs;
    auto add = [&](int parent, int jt, double mass, double len, double wid) {
        segs.push_back({parent, jt, mass, len, wid}); return (int)segs.size(); };
    const int pelvis = add(0, SBK_JOINT_FREE, 11.0, 0.20, 0.30);
    const int torso  = add(pelvis, SBK_JOINT_BALL, 14.0, 0.25, 0.28);
    const int chest  = add(torso, SBK_JOINT_UNIVERSAL, 30.0, 0.30, 0.32);
    const int neck   = add(chest, SBK_JOINT_BALL, 1.2, 0.10, 0.10);
    const int head   = add(neck, SBK_JOINT_PIN, 4.5, 0.22, 0.18);
    (void)            add(head, SBK_JOINT_PIN, 0.3, 0.08, 0.10);              // jaw
    for (int side = 0; side < 2; ++side) {
        int p = add(chest, SBK_JOINT_PIN, 0.4, 0.15, 0.04);                  // clavicle
        p = add(p, SBK_JOINT_BALL, 2.1, 0.30, 0.09);                          // shoulder/upper arm
        p = add(p, SBK_JOINT_PIN, 0.9, 0.14, 0.07);                           // elbow
        p = add(p, SBK_JOINT_PIN, 0.7, 0.13, 0.06);                           // forearm twist
        p = add(p, SBK_JOINT_UNIVERSAL, 0.4, 0.09, 0.08);                     // wrist/hand
        (void)add(p, SBK_JOINT_PIN, 0.05, 0.08, 0.07);                        // fingers
    }
    for (int side = 0; side < 2; ++side) {
        int p = add(pelvis, SBK_JOINT_BALL, 9.0, 0.42, 0.15);                 // hip/thigh
        p = add(p, SBK_JOINT_SLIDER, 0.1, 0.05, 0.05);                        // patella
        p = add(p, SBK_JOINT_PIN, 3.6, 0.43, 0.10);                           // knee/shank
        p = add(p, SBK_JOINT_UNIVERSAL, 0.2, 0.06, 0.07);                     // ankle/talus
        p = add(p, SBK_JOINT_PIN, 0.9, 0.17, 0.08);                           // subtalar/foot
        (void)add(p, SBK_JOINT_PIN, 0.15, 0.07, 0.08);                        // toes
    }
    for (size_t i = 0; i < segs.size(); ++i) {
        const Seg& s = segs[i];
        sbk_body_desc b = groundBody(); b.parent = s.parent; b.joint_type = s.jt; b.mass = s.mass;
        b.com_B[0] = 0.02*s.len*rng.next(); b.com_B[1] = -0.5*s.len + 0.05*s.len*rng.next(); b.com_B[2] = 0.02*s.len*rng.next();
        boxUnitInertia(b.unit_inertia_OB_B, s.wid, s.len, 0.8*s.wid, b.com_B,
                       0.3*rng.next(), rng.next(), rng.next(), rng.next() + 1.5);
        const double plen = s.parent > 0 ? segs[s.parent-1].len : 1.0;
        const double pwid = s.parent > 0 ? segs[s.parent-1].wid : 0.0;
        setAxisAngleX(b.X_PF, 0.6*rng.next(), rng.next(), rng.next(), rng.next() + 1.2,
                      0.5*pwid*rng.next(), (s.parent > 0 ? -plen : 1.0) + 0.1*plen*rng.next(), 0.3*pwid*rng.next());
        setAxisAngleX(b.X_BM, 0.5*rng.next(), rng.next() + 1.1, rng.next(), rng.next(),
                      0.1*s.wid*rng.next(), 0.05*s.len*rng.next(), 0.1*s.wid*rng.next());
        m.bodies.push_back(b);
    }
    m.forces.push_back(gravityForce(9.80665, 0, -1, 0));
    // Joint impedances scaled by the segment's own inertia so that every mobility has a natural
    // frequency of at most ~30 rad/s and light damping: a non-stiff system at h = 1e-3.
    for (int i = 1; i < (int)m.bodies.size(); ++i) {
        const int jt = m.bodies[i].joint_type; const Seg& sg = segs[i-1];
        const double Ieff = (jt == SBK_JOINT_SLIDER) ? sg.mass : sg.mass*sg.len*sg.len;
        if (jt == SBK_JOINT_PIN || jt == SBK_JOINT_SLIDER || jt == SBK_JOINT_UNIVERSAL) {
            for (int c = 0; c < jointNQ(jt); ++c) {
                m.forces.push_back(springForce(i, c, 900.0*Ieff, 0.1*rng.next()));
                m.forces.push_back(damperForce(i, c, 6.0*Ieff));
            }
        } else if (jt == SBK_JOINT_BALL) {
            for (int c = 0; c < 3; ++c) m.forces.push_back(damperForce(i, c, 6.0*Ieff));
        }
    }
    return m;
}

// C5: n-body branched tree, parent(i) = floor(i/2) (i>=2), body 1 on Ground; joint by i mod 4:
// 0 Ball, 1 Universal, 2/3 Pin; link offset (0,0.3,0); gravity.  (Pattern of the reference's
// Simbody/tests/adhoc/TestMultibodyPerformance.cpp:288-305.)
inline ModelSpec makeBranchedTree(int n) {
    ModelSpec m; m.name = "branched_tree"; m.bodies.push_back(groundBody());
    XorShift64 rng(0xD1B54A32D192ED03ull);
    for (int i = 1; i <= n; ++i) {
        sbk_body_desc b = groundBody();
        b.parent = (i == 1) ? 0 : i/2;
        const int r = i % 4;
        b.joint_type = r == 0 ? SBK_JOINT_BALL : r == 1 ? SBK_JOINT_UNIVERSAL : SBK_JOINT_PIN;
        b.mass = 0.5 + 0.25*(rng.next() + 1);
        b.com_B[0] = 0.01*rng.next(); b.com_B[1] = -0.15 + 0.01*rng.next(); b.com_B[2] = 0.01*rng.next();
        boxUnitInertia(b.unit_inertia_OB_B, 0.08, 0.3, 0.06, b.com_B, 0.2*rng.next(), rng.next(), rng.next(), rng.next() + 1.5);
        setAxisAngleX(b.X_PF, 0.7*rng.next(), rng.next(), rng.next(), rng.next() + 1.2, 0.02*rng.next(), -0.3, 0.02*rng.next());
        setAxisAngleX(b.X_BM, 0.4*rng.next(), rng.next() + 1.1, rng.next(), rng.next(), 0, 0, 0);
        m.bodies.push_back(b);
    }
    m.forces.push_back(gravityForce(9.80665, 0, -1, 0));
    return m;
}

// A small mixed tree driven by Force::UniformGravity (a gravity vector that is not axis aligned) and
// Force::GlobalDamper plus one spring: Ball -> Universal -> Pin -> Slider with a Pin branch off body 1.
inline ModelSpec makeUgDamp5() {
    ModelSpec m = makeMixed7(); m.name = "ugdamp5";
    m.bodies.erase(m.bodies.begin() + 1);                                   // drop the Free root
    for (size_t i = 1; i < m.bodies.size(); ++i) m.bodies[i].parent = (i == 5) ? 1 : (int)i - 1;
    m.forces.clear();
    m.forces.push_back(uniformGravityForce(0.3, -9.7, 0.4));
    m.forces.push_back(springForce(3, 0, 12.0, -0.1));
    m.forces.push_back(globalDamperForce(0.8));
    m.forces.push_back(mobilityConstantForce(2, 1, 1.7));      // a motor torque on the Universal's second speed
    return m;
}

// mixed7 with two Weld mobilizers: one in the middle of the chain (bodies keep moving outboard of
// it) and one as a leaf carrying extra mass.
inline ModelSpec makeWelded8() {
    ModelSpec m = makeMixed7(); m.name = "welded8";
    m.bodies[3].joint_type = SBK_JOINT_WELD;                 // was Universal: Free -> Ball -> WELD -> Pin -> Slider
    sbk_body_desc leaf = m.bodies[6]; leaf.parent = 4; leaf.joint_type = SBK_JOINT_WELD; leaf.mass = 0.7;
    m.bodies.push_back(leaf);
    m.forces.clear();
    m.forces.push_back(gravityForce(9.81, 0, -1, 0));
    m.forces.push_back(springForce(4, 0, 30.0, 0.2));
    m.forces.push_back(damperForce(5, 0, 1.5));
    return m;
}

// Planar -> Cylinder -> Translation -> Pin -> Gimbal chain with Slider and Cylinder branches: the axis-aligned mobilizers with
// Cartesian coordinates, general frames, a spring on a translational and one on a rotational coordinate.
inline ModelSpec makeCartesian8() {
    ModelSpec m = makeMixed7(); m.name = "cartesian8";
    m.bodies[1].joint_type = SBK_JOINT_PLANAR; m.bodies[2].joint_type = SBK_JOINT_CYLINDER;
    m.bodies[3].joint_type = SBK_JOINT_TRANSLATION; m.bodies[4].joint_type = SBK_JOINT_PIN; m.bodies[5].joint_type = SBK_JOINT_SLIDER;
    m.bodies[5].parent = 2;
    m.bodies[6].joint_type = SBK_JOINT_CYLINDER;
    sbk_body_desc g = m.bodies[4]; g.parent = 4; g.joint_type = SBK_JOINT_GIMBAL; m.bodies.push_back(g);   // body 7: Gimbal off the Pin
    m.forces.clear();
    m.forces.push_back(gravityForce(9.81, 0, -1, 0));
    m.forces.push_back(springForce(3, 1, 25.0, 0.1));
    m.forces.push_back(springForce(2, 0, 8.0, -0.2));
    m.forces.push_back(damperForce(1, 2, 0.7));
    return m;
}

// mixed7 plus the two-point elements of Force.cpp:103-221: a spring between stations on the Universal body and the branch Pin body, a
// spring from a Ground station to the Slider body, and a damper between the Ball body and the last Pin body (ExampleLongPendulum-style
// models attach such elements between bodies and Ground).
inline ModelSpec makeTwoPoint7() {
    ModelSpec m = makeMixed7(); m.name = "twopoint7";
    const double s1[3] = {0.1, -0.3, 0.05}, s2[3] = {-0.2, 0.15, 0.1}, g0[3] = {0.5, 1.5, -0.4}, s3[3] = {0.05, 0.1, -0.12}, s4[3] = {0.0, -0.2, 0.07};
    m.forces.push_back(twoPointSpringForce(3, s1, 6, s2, 40.0, 0.35));
    m.forces.push_back(twoPointSpringForce(0, g0, 5, s3, 15.0, 0.8));
    m.forces.push_back(twoPointDamperForce(2, s4, 4, s1, 2.5));
    return m;
}

inline ModelSpec makeNamedModel(const std::string& name, int n) {
    if (name == "double_pendulum") return makePinChain(2, "double_pendulum");
    if (name == "pin_chain")       return makePinChain(n > 0 ? n : 50);
    if (name == "mixed7")          return makeMixed7();
    if (name == "mixed7e")         { ModelSpec m = makeMixed7(); m.name = "mixed7e"; m.useEulerAngles = true; return m; }   // Euler-angle mode
    if (name == "ugdamp5")         return makeUgDamp5();
    if (name == "welded8")         return makeWelded8();
    if (name == "cartesian8")      return makeCartesian8();
    if (name == "twopoint7")       return makeTwoPoint7();
    if (name == "humanoid30")      return makeHumanoid30();
    if (name == "branched_tree")   return makeBranchedTree(n > 0 ? n : 1000);
    throw std::runtime_error("unknown model '" + name + "'");
}

// ---- text serialisation ----------------------------------------------------------------------
inline std::string toText(const ModelSpec& m) {
    std::string out; char buf[64];
    auto num = [&](double v) { std::snprintf(buf, sizeof buf, " %.17g", v); out += buf; };
    out += "sbkmodel 1\nname " + m.name + "\n" + (m.useEulerAngles ? "euler 1\n" : "") + "nb " + std::to_string(m.bodies.size()) + "\n";
    for (size_t i = 0; i < m.bodies.size(); ++i) {
        const sbk_body_desc& b = m.bodies[i];
        out += "body " + std::to_string(i) + " " + std::to_string(b.parent) + " " + jointName(b.joint_type);
        num(b.mass); for (double v : b.com_B) num(v); for (double v : b.unit_inertia_OB_B) num(v);
        for (double v : b.X_PF) num(v); for (double v : b.X_BM) num(v);
        out += "\n";
    }
    out += "nf " + std::to_string(m.forces.size()) + "\n";
    for (const sbk_force_desc& f : m.forces) {
        if (f.kind == SBK_FORCE_GRAVITY) { out += "gravity"; num(f.a); num(f.dir[0]); num(f.dir[1]); num(f.dir[2]); }
        else if (f.kind == SBK_FORCE_SPRING) { out += "spring " + std::to_string(f.body) + " " + std::to_string(f.coord); num(f.a); num(f.b); }
        else if (f.kind == SBK_FORCE_DAMPER) { out += "damper " + std::to_string(f.body) + " " + std::to_string(f.coord); num(f.a); }
        else if (f.kind == SBK_FORCE_UNIFORM_GRAVITY) { out += "ugravity"; num(f.dir[0]); num(f.dir[1]); num(f.dir[2]); }
        else if (f.kind == SBK_FORCE_GLOBAL_DAMPER) { out += "gdamper"; num(f.a); }
        else if (f.kind == SBK_FORCE_MOBILITY_CONSTANT) { out += "mconst " + std::to_string(f.body) + " " + std::to_string(f.coord); num(f.a); }
        else if (f.kind == SBK_FORCE_TWO_POINT_SPRING || f.kind == SBK_FORCE_TWO_POINT_DAMPER) {
            out += (f.kind == SBK_FORCE_TWO_POINT_SPRING ? "tpspring " : "tpdamper ") + std::to_string(f.body) + " " + std::to_string(f.coord);
            num(f.a); if (f.kind == SBK_FORCE_TWO_POINT_SPRING) num(f.b);
            for (double v : f.dir) num(v); for (double v : f.station2) num(v);
        }
        else throw std::runtime_error("bad force kind");
        out += "\n";
    }
    return out;
}

inline ModelSpec fromText(const std::string& text) {
    std::istringstream in(text);
    std::string tok; int ver = 0;
    if (!(in >> tok >> ver) || tok != "sbkmodel" || ver != 1) throw std::runtime_error("not an sbkmodel v1 text");
    ModelSpec m; int nb = 0, nf = 0;
    in >> tok >> m.name;  if (tok != "name") throw std::runtime_error("expected 'name'");
    in >> tok >> nb;
    if (tok == "euler") { m.useEulerAngles = nb != 0; in >> tok >> nb; }
    if (tok != "nb" || nb < 1) throw std::runtime_error("expected 'nb'");
    for (int i = 0; i < nb; ++i) {
        int idx; std::string jn; sbk_body_desc b; std::memset(&b, 0, sizeof b);
        in >> tok >> idx >> b.parent >> jn;
        if (!in || tok != "body" || idx != i) throw std::runtime_error("bad body line " + std::to_string(i));
        b.joint_type = jointFromName(jn);
        in >> b.mass; for (double& v : b.com_B) in >> v; for (double& v : b.unit_inertia_OB_B) in >> v;
        for (double& v : b.X_PF) in >> v; for (double& v : b.X_BM) in >> v;
        if (!in) throw std::runtime_error("truncated body line " + std::to_string(i));
        m.bodies.push_back(b);
    }
    in >> tok >> nf; if (!in || tok != "nf") throw std::runtime_error("expected 'nf'");
    for (int i = 0; i < nf; ++i) {
        sbk_force_desc f; std::memset(&f, 0, sizeof f);
        in >> tok;
        if (tok == "gravity") { f.kind = SBK_FORCE_GRAVITY; f.body = -1; in >> f.a >> f.dir[0] >> f.dir[1] >> f.dir[2]; }
        else if (tok == "spring") { f.kind = SBK_FORCE_SPRING; in >> f.body >> f.coord >> f.a >> f.b; }
        else if (tok == "damper") { f.kind = SBK_FORCE_DAMPER; in >> f.body >> f.coord >> f.a; }
        else if (tok == "ugravity") { f.kind = SBK_FORCE_UNIFORM_GRAVITY; f.body = -1; f.a = 1; in >> f.dir[0] >> f.dir[1] >> f.dir[2]; }
        else if (tok == "gdamper") { f.kind = SBK_FORCE_GLOBAL_DAMPER; f.body = -1; in >> f.a; }
        else if (tok == "mconst") { f.kind = SBK_FORCE_MOBILITY_CONSTANT; in >> f.body >> f.coord >> f.a; }
        else if (tok == "tpspring" || tok == "tpdamper") {
            f.kind = tok == "tpspring" ? SBK_FORCE_TWO_POINT_SPRING : SBK_FORCE_TWO_POINT_DAMPER;
            in >> f.body >> f.coord >> f.a; if (tok == "tpspring") in >> f.b;
            for (double& v : f.dir) in >> v; for (double& v : f.station2) in >> v;
        }
        else throw std::runtime_error("bad force token " + tok);
        if (!in) throw std::runtime_error("truncated force line");
        m.forces.push_back(f);
    }
    return m;
}

} // namespace sbk
