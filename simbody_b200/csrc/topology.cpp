// topology.cpp -- see topology.h
#include "topology.h"
#include <algorithm>
#include <cstring>

namespace sbk {

void compileTopology(const ModelSpec& spec, sbk_topology& t) {
    const int nb = (int)spec.bodies.size();
    if (nb < 1) throw std::runtime_error("topology: no bodies (entry 0 must be Ground)");
    if (spec.bodies[0].joint_type != SBK_JOINT_GROUND || spec.bodies[0].parent != -1)
        throw std::runtime_error("topology: body 0 must be Ground with parent -1");
    t.spec = spec; t.nb = nb;
    t.q0.assign(nb, 0); t.nqOf.assign(nb, 0); t.u0.assign(nb, 0); t.nuOf.assign(nb, 0);
    t.level.assign(nb, 0); t.quatIndex.assign(nb, -1);
    t.bodies.assign(nb, sbkd::BodyConst());
    std::vector<std::vector<int>> kids(nb);
    int nextQ = 0, nextU = 0, nextQuat = 0; bool chain = true;
    for (int b = 1; b < nb; ++b) {
        const sbk_body_desc& d = spec.bodies[b];
        if (d.parent < 0 || d.parent >= b)
            throw std::runtime_error("topology: body " + std::to_string(b) + " has parent " + std::to_string(d.parent) +
                                     "; a tree needs 0 <= parent < body index");
        if (d.joint_type < SBK_JOINT_PIN || d.joint_type > SBK_JOINT_GIMBAL)
            throw std::runtime_error("topology: body " + std::to_string(b) + " has unsupported mobilizer kind " + std::to_string(d.joint_type));
        if (!(d.mass > 0)) throw std::runtime_error("topology: body " + std::to_string(b) + " needs mass > 0");
        if (d.parent != b - 1) chain = false;
        t.level[b] = t.level[d.parent] + 1;
        t.q0[b] = nextQ; t.nqOf[b] = jointNQ(d.joint_type); nextQ += t.nqOf[b];
        t.u0[b] = nextU; t.nuOf[b] = jointNU(d.joint_type); nextU += t.nuOf[b];
        if ((d.joint_type == SBK_JOINT_BALL || d.joint_type == SBK_JOINT_FREE) && !spec.useEulerAngles) t.quatIndex[b] = nextQuat++;
        kids[d.parent].push_back(b);
    }
    t.nq = nextQ; t.nu = nextU; t.nquat = nextQuat; t.isChain = chain && nb > 1;
    t.nlevels = 1 + *std::max_element(t.level.begin(), t.level.end());

    // level order
    t.levelOrder.resize(nb);
    for (int b = 0; b < nb; ++b) t.levelOrder[b] = b;
    std::stable_sort(t.levelOrder.begin(), t.levelOrder.end(), [&](int a, int b) { return t.level[a] < t.level[b]; });
    t.levelStart.assign(t.nlevels + 1, 0);
    for (int b = 0; b < nb; ++b) t.levelStart[t.level[b] + 1]++;
    for (int l = 0; l < t.nlevels; ++l) t.levelStart[l+1] += t.levelStart[l];
    t.maxLevelWidth = 0;
    for (int l = 0; l < t.nlevels; ++l) t.maxLevelWidth = std::max(t.maxLevelWidth, t.levelStart[l+1] - t.levelStart[l]);

    // forces
    int ngrav = 0;
    std::vector<std::vector<sbkd::ForceConst>> perBody(nb);
    for (size_t i = 0; i < spec.forces.size(); ++i) {
        const sbk_force_desc& f = spec.forces[i];
        if (f.kind == SBK_FORCE_GRAVITY) {
            if (++ngrav > 1) throw std::runtime_error("topology: more than one gravity element is not supported");
            // Force_Gravity.cpp:532: gravity = g * d
            t.grav[0] = f.a*f.dir[0]; t.grav[1] = f.a*f.dir[1]; t.grav[2] = f.a*f.dir[2];
        } else if (f.kind == SBK_FORCE_UNIFORM_GRAVITY) {
            if (++ngrav > 1) throw std::runtime_error("topology: more than one gravity element is not supported");
            // Force.cpp:1053: frc_G = m*g with the user's vector g
            t.grav[0] = f.dir[0]; t.grav[1] = f.dir[1]; t.grav[2] = f.dir[2];
        } else if (f.kind == SBK_FORCE_GLOBAL_DAMPER) {
            // Force.cpp:997: mobilityForces -= damping*u, i.e. a linear damper on every mobility, at this force index
            for (int b = 1; b < nb; ++b)
                for (int j = 0; j < jointNU(spec.bodies[b].joint_type); ++j) {
                    sbkd::ForceConst fc; fc.kind = SBK_FORCE_DAMPER; fc.coord = j; fc.a = f.a; fc.b = 0;
                    perBody[b].push_back(fc);
                }
        } else if (f.kind == SBK_FORCE_SPRING || f.kind == SBK_FORCE_DAMPER || f.kind == SBK_FORCE_MOBILITY_CONSTANT) {
            if (f.body < 1 || f.body >= nb) throw std::runtime_error("topology: force " + std::to_string(i) + " acts on invalid body");
            const int jt = spec.bodies[f.body].joint_type;
            if (f.coord < 0 || f.coord >= jointNU(jt)) throw std::runtime_error("topology: force " + std::to_string(i) + " has invalid coordinate");
            if (f.kind == SBK_FORCE_SPRING && (jt == SBK_JOINT_BALL || jt == SBK_JOINT_FREE))
                throw std::runtime_error("topology: MobilityLinearSpring on Ball/Free is not supported (reference Force.cpp:348 mixes q/u indices)");
            sbkd::ForceConst fc; fc.kind = f.kind; fc.coord = f.coord; fc.a = f.a; fc.b = f.b;
            perBody[f.body].push_back(fc);
        } else throw std::runtime_error("topology: unknown force kind");
    }

    t.children.clear(); t.forces.clear();
    for (int b = 0; b < nb; ++b) {
        const sbk_body_desc& d = spec.bodies[b];
        sbkd::BodyConst& bc = t.bodies[b];
        std::memset(&bc, 0, sizeof bc);
        std::memcpy(bc.X_PF, d.X_PF, sizeof bc.X_PF);
        // X_MB = ~X_BM = (R^T, -(R^T p))
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) bc.X_MB[3*i+j] = d.X_BM[3*j+i];
        for (int i = 0; i < 3; ++i)
            bc.X_MB[9+i] = -(bc.X_MB[3*i]*d.X_BM[9] + bc.X_MB[3*i+1]*d.X_BM[10] + bc.X_MB[3*i+2]*d.X_BM[11]);
        bc.mass = d.mass;
        for (int i = 0; i < 3; ++i) bc.p_BM[i] = d.X_BM[9+i];
        for (int i = 0; i < 3; ++i) bc.com_B[i] = d.com_B[i];
        for (int i = 0; i < 6; ++i) bc.G_B[i] = d.unit_inertia_OB_B[i];
        bc.joint = d.joint_type; bc.parent = d.parent < 0 ? 0 : d.parent;
        if (spec.useEulerAngles && d.joint_type == SBK_JOINT_BALL) bc.joint = sbkd::JT_BALL_EULER;     // internal kinds: x-y-z angles
        if (spec.useEulerAngles && d.joint_type == SBK_JOINT_FREE) bc.joint = sbkd::JT_FREE_EULER;
        bc.q0 = t.q0[b]; bc.u0 = t.u0[b]; bc.quat = t.quatIndex[b]; bc.level = t.level[b];
        bc.flags = 0;
        auto isIdentityR = [](const double* X) { return X[0] == 1 && X[4] == 1 && X[8] == 1 && X[1] == 0 && X[2] == 0 &&
                                                        X[3] == 0 && X[5] == 0 && X[6] == 0 && X[7] == 0; };
        if (isIdentityR(bc.X_PF)) bc.flags |= sbkd::BF_NO_R_PF;
        if (isIdentityR(bc.X_MB)) bc.flags |= sbkd::BF_NO_R_MB;
        if (b >= 1 && d.parent == b - 1) bc.flags |= sbkd::BF_PARENT_PREV;
        for (int k : kids[b]) if (k != b + 1) bc.flags |= sbkd::BF_STORE_LINK;
        if (b >= 1 && (kids[b].empty() || kids[b][0] != b + 1)) bc.flags |= sbkd::BF_TIP;
        bc.nchild = (int)kids[b].size(); bc.childStart = (int)t.children.size();
        for (int k : kids[b]) t.children.push_back(k);
        bc.nforce = (int)perBody[b].size(); bc.forceStart = (int)t.forces.size();
        for (const auto& fc : perBody[b]) t.forces.push_back(fc);
    }
    if (t.children.empty()) t.children.push_back(0);
    if (t.forces.empty()) { sbkd::ForceConst z; std::memset(&z, 0, sizeof z); t.forces.push_back(z); }
}

} // namespace sbk
