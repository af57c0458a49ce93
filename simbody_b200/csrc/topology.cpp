// topology.cpp -- see topology.h
#include "topology.h"
#include "sbk_ltree.cuh"
#include <algorithm>
#include <cstring>

namespace sbk {

// Constants of the body-frame integrator path (sbk_local.cuh): every body's quantities are expressed in its own
// outboard frame M.  X_T = X_MB(parent) * X_PF takes the child's inboard frame F to the parent's M; the mass
// properties move from (Bo, B) to (Mo, M) once, here.
static void compileLocalTables(sbk_topology& t) {
    const int nb = t.nb;
    t.localOk = nb > 1 && t.twoPoint.empty(); t.lbodies.clear(); t.lrows = 0;     // two-point elements need ground-frame positions
    for (int b = 1; b < nb; ++b) {
        const int jt = t.bodies[b].joint;
        if (jt != sbkd::JT_PIN && jt != sbkd::JT_SLIDER && jt != sbkd::JT_UNIVERSAL && jt != sbkd::JT_BALL && jt != sbkd::JT_FREE) t.localOk = false;
    }
    if (!t.localOk) return;
    t.lbodies.assign(nb, sbkd::LBody());
    int row = 0;
    for (int b = 0; b < nb; ++b) {
        const sbkd::BodyConst& bc = t.bodies[b];
        sbkd::LBody& lb = t.lbodies[b];
        std::memset(&lb, 0, sizeof lb);
        const double* Rmb = bc.X_MB; const double* pmb = bc.X_MB + 9;              // this body's X_MB (B -> M)
        // parent's X_MB (identity for Ground and for b == 0)
        double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pp[3] = {0, 0, 0};
        if (b > 0 && bc.parent > 0) { std::memcpy(Rp, t.bodies[bc.parent].X_MB, sizeof Rp); std::memcpy(pp, t.bodies[bc.parent].X_MB + 9, sizeof pp); }
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) lb.RT[3*i+j] = Rp[3*i]*bc.X_PF[j] + Rp[3*i+1]*bc.X_PF[3+j] + Rp[3*i+2]*bc.X_PF[6+j];
            lb.pT[i] = pp[i] + (Rp[3*i]*bc.X_PF[9] + Rp[3*i+1]*bc.X_PF[10] + Rp[3*i+2]*bc.X_PF[11]);
        }
        // mass properties about Mo in M
        const double m = bc.mass; lb.m = m;
        double cM[3], dB[3];                                                    // com from Mo, in M and in B
        for (int i = 0; i < 3; ++i) cM[i] = Rmb[3*i]*bc.com_B[0] + Rmb[3*i+1]*bc.com_B[1] + Rmb[3*i+2]*bc.com_B[2] + pmb[i];
        for (int i = 0; i < 3; ++i) { dB[i] = bc.com_B[i] - bc.p_BM[i]; lb.h[i] = m*cM[i]; }
        const double* cB = bc.com_B;
        double IB[3][3] = {{bc.G_B[0], bc.G_B[3], bc.G_B[4]}, {bc.G_B[3], bc.G_B[1], bc.G_B[5]}, {bc.G_B[4], bc.G_B[5], bc.G_B[2]}};   // unit inertia about Bo
        const double c2 = cB[0]*cB[0] + cB[1]*cB[1] + cB[2]*cB[2], d2 = dB[0]*dB[0] + dB[1]*dB[1] + dB[2]*dB[2];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)                 // to the mass centre, then out to Mo (per unit mass, B axes)
            IB[i][j] += -((i == j ? c2 : 0) - cB[i]*cB[j]) + ((i == j ? d2 : 0) - dB[i]*dB[j]);
        double IM[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) s += Rmb[3*i+k]*IB[k][l]*Rmb[3*j+l];
            IM[i][j] = m*s;
        }
        lb.I[0] = IM[0][0]; lb.I[1] = IM[1][1]; lb.I[2] = IM[2][2];
        lb.I[3] = 0.5*(IM[0][1] + IM[1][0]); lb.I[4] = 0.5*(IM[0][2] + IM[2][0]); lb.I[5] = 0.5*(IM[1][2] + IM[2][1]);
        lb.joint = bc.joint; lb.parent = bc.parent; lb.q0 = bc.q0; lb.u0 = bc.u0;
        lb.flags = bc.flags & (sbkd::BF_PARENT_PREV | sbkd::BF_STORE_LINK | sbkd::BF_TIP);
        bool idRT = true;
        for (int i = 0; i < 9; ++i) if (lb.RT[i] != ((i % 4 == 0) ? 1.0 : 0.0)) idRT = false;
        if (idRT) lb.flags |= sbkd::BF_NO_RT;
        lb.nchild = bc.nchild; lb.childStart = bc.childStart; lb.nforce = bc.nforce; lb.forceStart = bc.forceStart;
        lb.rec = row; row += sbkd::lrSize(t.nuOf[b]);
    }
    for (int b = 1; b < nb; ++b) { const int p = t.lbodies[b].parent; t.lbodies[b].parentLink = t.lbodies[p].rec + sbkd::lrV(t.nuOf[p]); }
    for (int b = 0; b < nb; ++b)
        for (int j = 0; j < 4; ++j) {
            const sbkd::LBody& lb = t.lbodies[b];
            const int cidx = j < lb.nchild ? t.children[lb.childStart + j] : -1;
            t.lbodies[b].childIA[j] = cidx < 0 ? -1 : t.lbodies[cidx].rec + sbkd::lrIA(t.nuOf[cidx]);
        }
    t.lrows = row;
    // level-order variant: no body finds its parent in a carry (every link through the records), every body keeps its own
    // velocity for the inward sweep (BF_TIP), every body with children publishes v / a (BF_STORE_LINK)
    t.lbodiesLevel = t.lbodies;
    for (int b = 0; b < nb; ++b) {
        int f = t.lbodiesLevel[b].flags & sbkd::BF_NO_RT;
        if (b >= 1) f |= sbkd::BF_TIP;
        if (t.lbodiesLevel[b].nchild > 0) f |= sbkd::BF_STORE_LINK;
        t.lbodiesLevel[b].flags = f;
    }
    t.lfcoef.assign((size_t)3*std::max(1, t.nu), 0.0);
    for (int b = 1; b < nb; ++b) {
        const sbkd::BodyConst& bc = t.bodies[b];
        for (int k = 0; k < bc.nforce; ++k) {
            const sbkd::ForceConst& fc = t.forces[bc.forceStart + k];
            double* c = &t.lfcoef[(size_t)3*(bc.u0 + fc.coord)];
            if (fc.kind == sbkd::FK_SPRING) { c[0] += fc.a*fc.b; c[1] -= fc.a; }
            else if (fc.kind == sbkd::FK_CONSTANT) c[0] += fc.a;
            else c[2] -= fc.a;
        }
    }
}

// nclusters > 1: the nwarps warps belong to that many clusters (consecutive ranges); the levels above the cut run on the FIRST
// cluster's warps (cluster barrier between levels) and the two parts of a sweep meet at a barrier of all the clusters (LT_XSYNC).
// CTA-local levels: with one subtree per warp, a body above the cut whose descendants' subtrees all belong to the eight warps of
// ONE CTA is processed by the warp of its first child, right after / before that warp's subtree, with the CTA's barrier
// (LT_TSYNC, every warp of every CTA one entry per such level) -- for a binary tree the three levels above the cut leave the
// cluster-barrier part of the sweep.
TreeCut cutTreeForWarps(const sbk_topology& t, int nwarps, int topWarps, int cutWidth, int nclusters) {
    if (cutWidth <= 0) cutWidth = nwarps;
    TreeCut r; r.bodies = t.lbodiesLevel; r.cutLevel = t.nlevels; r.subStart.assign(1, 0);
    for (int l = 1; l < t.nlevels; ++l) if (t.levelStart[l + 1] - t.levelStart[l] >= cutWidth) { r.cutLevel = l; break; }
    const int L = r.cutLevel;
    std::vector<std::vector<int>> dfs;
    for (int i = t.levelStart[std::min(L, t.nlevels - 1)]; L < t.nlevels && i < t.levelStart[L + 1]; ++i) {
        // depth-first walk, children in list order (the first child directly follows its parent)
        dfs.emplace_back();
        std::vector<int> stack(1, t.levelOrder[i]);
        while (!stack.empty()) {
            const int b = stack.back(); stack.pop_back();
            dfs.back().push_back(b);
            const sbkd::LBody& lb = t.lbodies[b];
            for (int j = lb.nchild - 1; j >= 0; --j) stack.push_back(t.children[lb.childStart + j]);
        }
    }
    const int nsub = (int)dfs.size();
    const int perCluster = nwarps/std::max(1, nclusters);
    topWarps = std::max(1, std::min(topWarps, nclusters > 1 ? perCluster : nwarps));
    const bool multiCta = topWarps > 8 || nclusters > 1;
    // CTA-local levels l0 .. L-1 (l0 = L: none) and the chain of such bodies each subtree's warp owns (top-down)
    int l0 = L;
    std::vector<std::vector<int>> chain(nsub);
    if (nsub > 0 && nsub <= nwarps && multiCta) {
        std::vector<int> owner(t.nb, -1);
        for (int s = 0; s < nsub; ++s) owner[dfs[s][0]] = s;
        for (int l = L - 1; l >= 1; --l) {
            std::vector<std::pair<int, int>> got;
            bool all = true;
            for (int i = t.levelStart[l]; all && i < t.levelStart[l + 1]; ++i) {
                const int p = t.levelOrder[i]; const sbkd::LBody& lb = t.lbodies[p];
                if (lb.nchild == 0) { all = false; break; }
                int lo = nwarps, hi = -1;
                for (int j = 0; j < lb.nchild; ++j) { const int o = owner[t.children[lb.childStart + j]]; if (o < 0) all = false; lo = std::min(lo, o); hi = std::max(hi, o); }
                if (!all || lo/8 != hi/8) { all = false; break; }
                got.push_back({p, owner[t.children[lb.childStart]]});
            }
            if (!all) break;
            for (const auto& g : got) owner[g.first] = g.second;
            l0 = l;
        }
        for (int l = l0; l < L; ++l)
            for (int i = t.levelStart[l]; i < t.levelStart[l + 1]; ++i) chain[owner[t.levelOrder[i]]].push_back(t.levelOrder[i]);
    }
    // the levels that remain on top: on the first CTA's eight warps with __syncthreads when none of them is wider than that,
    // else on topWarps warps of the first cluster with the cluster barrier
    int topWidth = 0;
    for (int l = 1; l < l0; ++l) topWidth = std::max(topWidth, t.levelStart[l + 1] - t.levelStart[l]);
    if (multiCta && l0 < L && topWidth <= 8) topWarps = std::min(topWarps, 8);
    const int levelSync = (topWarps > 8 || (nclusters > 1 && !(l0 < L && topWidth <= 8))) ? sbkd::LT_GSYNC : sbkd::LT_TSYNC;
    // a warp's walk: its chain (a parent -> first child sequence ending right above the subtree's root), then the subtree
    for (int s = 0; s < nsub; ++s) {
        r.subOrder.insert(r.subOrder.end(), chain[s].begin(), chain[s].end());
        r.subOrder.insert(r.subOrder.end(), dfs[s].begin(), dfs[s].end());
        r.subStart.push_back((int)r.subOrder.size());
    }
    // link flags of the walk: parent previous / next is my first child / some child is not next
    for (int s = 0; s < nsub; ++s)
        for (int k = r.subStart[s]; k < r.subStart[s + 1]; ++k) {
            const int b = r.subOrder[k]; sbkd::LBody& lb = r.bodies[b];
            const int prev = k > r.subStart[s] ? r.subOrder[k - 1] : -1, next = k + 1 < r.subStart[s + 1] ? r.subOrder[k + 1] : -1;
            int f = lb.flags & sbkd::BF_NO_RT;
            if (prev >= 0 && lb.parent == prev) f |= sbkd::BF_PARENT_PREV;
            const bool nextIsChild = next >= 0 && lb.nchild > 0 && t.children[lb.childStart] == next;
            if (!nextIsChild) f |= sbkd::BF_TIP;
            if (lb.nchild > (nextIsChild ? 1 : 0)) f |= sbkd::BF_STORE_LINK;
            lb.flags = f;
        }
    // task lists (sbk_ltree.cuh): per warp, inward then outward
    r.listStart.assign(2*(size_t)nwarps, 0);
    const bool haveTop = l0 > 1, haveSub = nsub > 0;
    for (int dir = 0; dir < 2; ++dir)
        for (int w = 0; w < nwarps; ++w) {
            r.listStart[(size_t)dir*nwarps + w] = (int)r.lists.size();
            std::vector<int> sub, top;
            if (l0 < L) {        // one subtree per warp, CTA-local levels between it and the top
                const int s = w < nsub ? w : -1, nch = s >= 0 ? (int)chain[s].size() : 0;
                auto levelBody = [&](int l) { return (s >= 0 && l >= L - nch) ? chain[s][l - (L - nch)] : 0; };
                if (dir) {
                    for (int l = l0; l < L; ++l) sub.push_back(levelBody(l) | sbkd::LT_TSYNC);
                    if (s >= 0) sub.insert(sub.end(), dfs[s].begin(), dfs[s].end());
                } else {
                    if (s >= 0) sub.insert(sub.end(), dfs[s].rbegin(), dfs[s].rend());
                    if (sub.empty()) sub.push_back(0);
                    sub.back() |= sbkd::LT_TSYNC;
                    for (int l = L - 1; l >= l0; --l) sub.push_back(levelBody(l) | (l > l0 ? sbkd::LT_TSYNC : 0));
                }
            } else
                for (int s = w; s < nsub; s += nwarps) {
                    if (dir) sub.insert(sub.end(), dfs[s].begin(), dfs[s].end());
                    else     sub.insert(sub.end(), dfs[s].rbegin(), dfs[s].rend());
                }
            const bool inTopCluster = nclusters <= 1 || w < perCluster;
            if ((w < topWarps || levelSync == sbkd::LT_GSYNC) && haveTop && inTopCluster)
                for (int l = dir ? 1 : l0 - 1; dir ? l < l0 : l >= 1; l += dir ? 1 : -1) {
                    const size_t before = top.size();
                    if (w < topWarps) for (int i = t.levelStart[l] + w; i < t.levelStart[l + 1]; i += topWarps) top.push_back(t.levelOrder[i]);
                    if (top.size() == before) top.push_back(0);            // no body of this level for this warp: barrier only
                    top.back() |= levelSync;
                }
            // a barrier separates the two parts: inward subtrees | top, outward top | subtrees
            auto& first = dir ? top : sub; auto& second = dir ? sub : top;
            if (nclusters > 1) {
                if (haveTop && haveSub) { if (first.empty()) first.push_back(0); first.back() = (first.back() & ~sbkd::LT_GSYNC) | sbkd::LT_XSYNC; }
            }
            // (one cluster with group barriers between the top levels: the last / first of them already separates the parts)
            else if (haveTop && haveSub && !(levelSync == sbkd::LT_GSYNC && dir == 1)) { if (first.empty()) first.push_back(0); first.back() |= sbkd::LT_GSYNC; }
            r.lists.insert(r.lists.end(), first.begin(), first.end());
            r.lists.insert(r.lists.end(), second.begin(), second.end());
            r.lists.push_back(sbkd::LT_END); r.lists.push_back(sbkd::LT_END); r.lists.push_back(sbkd::LT_END);
        }
    return r;
}

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
    int ngrav = 0; t.twoPoint.clear();
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
        } else if (f.kind == SBK_FORCE_TWO_POINT_SPRING || f.kind == SBK_FORCE_TWO_POINT_DAMPER) {
            // body = body1, coord = body2 (Ground allowed: Force.cpp:103-140 takes any two mobilized bodies)
            if (f.body < 0 || f.body >= nb || f.coord < 0 || f.coord >= nb || f.body == f.coord)
                throw std::runtime_error("topology: force " + std::to_string(i) + " needs two different valid bodies");
            sbkd::TwoPointConst tp; std::memset(&tp, 0, sizeof tp);
            tp.kind = f.kind; tp.body1 = f.body; tp.body2 = f.coord; tp.a = f.a; tp.b = f.b;
            for (int k = 0; k < 3; ++k) { tp.s1[k] = f.dir[k]; tp.s2[k] = f.station2[k]; }
            t.twoPoint.push_back(tp);
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
    compileLocalTables(t);
    if (t.children.empty()) t.children.push_back(0);
    if (t.forces.empty()) { sbkd::ForceConst z; std::memset(&z, 0, sizeof z); t.forces.push_back(z); }
}

} // namespace sbk
