// topology.h -- host-side topology compiler: ModelSpec -> flat body table.
//
// Replaces the parts of SimbodyMatterSubsystemRep::endConstruction
// (Simbody/src/SimbodyMatterSubsystemRep.cpp:256-330) that the hot path needs:
//   * q/u slot assignment in MobilizedBodyIndex order using max-nq per mobilizer
//     (RigidBodyNodeSpec.h:81-87, SimbodyMatterSubsystemRep.cpp:270-293,537-571);
//   * level of each body and level-ordered body lists (rbNodeLevels,
//     SimbodyMatterSubsystemRep.h:1366);
//   * children lists in creation order (RigidBodyNode::children);
//   * X_MB = ~X_BM precomputed once (RigidBodyNode.h:1012);
//   * per-body lists of mobility force elements in force-index order.
#pragma once
#include <string>
#include <vector>
#include "sbk.h"
#include "sbk_sweeps.cuh"
#include "sbk_local.cuh"
#include "../host/model_spec.h"

struct sbk_topology {
    sbk::ModelSpec spec;
    int nb = 0, nq = 0, nu = 0, nquat = 0, nlevels = 0;
    std::vector<int> q0, nqOf, u0, nuOf, level, quatIndex;
    std::vector<sbkd::BodyConst>  bodies;     // cacheBase fields filled per batch/plan
    std::vector<int>              children;   // concatenated child lists
    std::vector<sbkd::ForceConst> forces;     // concatenated per-body mobility force lists
    std::vector<sbkd::LBody>      lbodies;    // body-frame integrator path (sbk_local.cuh); empty unless localOk
    std::vector<sbkd::LBody>      lbodiesLevel; // the same with level-order link flags (sbk_ltree.cuh): every link through the records
    std::vector<double>           lfcoef;     // [nu][3]: tau_j = A + B*q_j + C*u_j of the lowered mobility forces (sbk_local.cuh)
    int                           lrows = 0;  // scratch rows per instance of the body-frame path
    bool                          localOk = false;   // every mobilizer is Pin / Slider / Universal / Ball / Free (quaternion mode)
    std::vector<sbkd::TwoPointConst> twoPoint; // Force::TwoPointLinearSpring / Damper elements in force-index order (cacheBase filled per plan)
    std::vector<int>              levelOrder; // body indices sorted by level (stable)
    std::vector<int>              levelStart; // nlevels+1 offsets into levelOrder
    double grav[3] = {0, 0, 0};
    int maxLevelWidth = 0;
    bool isChain = false;                      // every body's parent is the previous body
};

namespace sbk {
// Throws std::runtime_error with a message on invalid input.
void compileTopology(const ModelSpec& spec, sbk_topology& out);
// Plan 5 (sbk_ltree.cuh): cut the tree at the first level at least `nwarps` bodies wide (levels above it run level-parallel, every
// body at it roots a subtree that one warp walks depth-first).  Fills the walk order of the subtrees, their start offsets, the cut
// level (== nlevels if the tree is never that wide: no subtrees) and the body table with the link flags of that schedule.
// lists / listStart: the per-warp task lists of sbk_ltree.cuh (listStart[dir*nwarps + w], dir 0 inward, 1 outward); the first
// topWarps warps also run the levels above the cut.
struct TreeCut { int cutLevel = 0; std::vector<int> subOrder, subStart, lists, listStart; std::vector<sbkd::LBody> bodies; };
TreeCut cutTreeForWarps(const sbk_topology& t, int nwarps, int topWarps, int cutWidth = 0, int nclusters = 1);   // cutWidth: minimum width of the cut level (default nwarps)
}
