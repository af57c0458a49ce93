// sbk_ltree.cuh -- the fused body-frame integrator on a TASK-LIST schedule, for wide trees in small batches.
//
// Same body steps as sbk_lrkm.cuh (lInwardBody, lFusedOutBody).  A group of nw warps owns 32 instances (lanes = instances);
// the host cuts the tree at the first level with at least one body per warp (topology.cpp: cutTreeForWarps) and writes, per
// warp and sweep direction, the list of bodies that warp processes:
//   * every body of the cut level roots a subtree that ONE warp walks depth-first with no barrier (links to the previous /
//     next body of the walk ride in the warp's carry, like the thread-per-instance plan);
//   * the few levels above the cut run level-parallel on the first `topNw` warps of the group, with their own (cheap)
//     barrier after each level; every link there goes through the scratch records;
//   * the two parts meet at the group's barrier once per sweep.
// A body is always processed by the same warp, so its own rows (sin/cos, G / nu, velocity, state slots) stay thread-private
// and are prefetched across the barriers; only the children's P+ / z+ and the parent's v / a are read after a barrier.
// The device loop is a three-deep software pipeline over the list: body constants of task i+2 (cp.async into a per-warp
// shared-memory slot), rows of task i+1, arithmetic of task i.
// The reference's analogue of the level lists: rbNodeLevels, SimbodyMatterSubsystemRep.cpp:1133-1135.
#pragma once
#include "sbk_lrkm.cuh"

namespace sbkd {

// list entry: body index (0 = no body, only the barriers) | flags
// LT_TSYNC: barrier of the warp's CTA; LT_GSYNC: barrier of the warp's cluster; LT_XSYNC: barrier of ALL the clusters that share the
// group of instances (a group may be spread over several clusters: sbk_ctree.cu)
enum { LT_BODY_MASK = 0x00ffffff, LT_XSYNC = 1 << 27, LT_TSYNC = 1 << 28, LT_GSYNC = 1 << 29, LT_END = 1 << 30 };
struct LTaskLists { const int* entries; const int* start; };      // start[dir*nw + w] .. : entries of warp w, dir 0 = inward, 1 = outward; LT_END terminated

enum { LT_BODY_SLOTS = 3 };
// per-warp shared-memory slots for the body constants of the tasks in flight
struct LBodySlots { LBody* slot; };                                 // LT_BODY_SLOTS consecutive LBody

SBK_HD void lcopyBody(LBody* dst, const LBody* src, const int lane) {
#if defined(__CUDA_ARCH__)
    constexpr int CH = (int)(sizeof(LBody)/16);
    if (lane < CH)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(reinterpret_cast<char*>(dst) + 16*lane)),
                     "l"(reinterpret_cast<const char*>(src) + 16*lane) : "memory");
#else
    (void)lane; *dst = *src;
#endif
}
SBK_HD void lwarpSync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

#define SBK_LT_PFSLOT(par) (cy + (LF_PF + LPF_ROWS*((par) & 1))*SBK_CARRY_STRIDE)

// One sweep of one warp over its task list.  OUT = false: inward (articulated inertias), true: fused outward.
template <int JMASK, bool OUT, class TSYNC, class GSYNC>
SBK_HD void lListSweep(const Ctx& c0, const LTables& T, const int* lst, const LBodySlots& B, const int inst, const bool active, const int lane, double* cy,
                       const LRkmWork& w, const double* S, const LStage& sg, const int vr, const int vw, int& par, TSYNC topSync, GSYNC groupSync) {
    constexpr bool BLK = SBK_DEV_BLK;
    Ctx c = c0; c.q = S; c.u = S + (BLK ? (long long)c.nq*BLK_LANES : (long long)c.nq*c.sStride);
    auto bodyOf = [](const int e) { return (e & LT_END) ? 0 : (e & LT_BODY_MASK); };
    auto rows = [&](const LBody& nx, double* pf) {
        if constexpr (OUT) lPrefetchOut<JMASK, BLK>(c, nx, inst, pf, w, S, sg);
        else lPrefetchIn<JMASK, BLK>(c, nx, inst, pf, S, vr);
    };
    // prologue: constants of tasks 0 and 1, rows of task 0
    int e0 = lst[0], e1 = (e0 & LT_END) ? e0 : lst[1], e2 = (e1 & LT_END) ? e1 : lst[2];
    int k = 0;                                                     // position of e0 in the list
    lcopyBody(B.slot + 0, T.bodies + bodyOf(e0), lane); lpfCommit();
    lcopyBody(B.slot + 1, T.bodies + bodyOf(e1), lane); lpfCommit();
    lpfWait(); lwarpSync();
    if (bodyOf(e0) && active) rows(B.slot[0], SBK_LT_PFSLOT(par));
    lpfCommit();
#pragma unroll 1
    while (!(e0 & LT_END)) {
        lcopyBody(B.slot + (k + 2) % LT_BODY_SLOTS, T.bodies + bodyOf(e2), lane); lpfCommit();
        lpfWait(); lwarpSync();                                    // constants of task k+1 and rows of task k have landed
        if (bodyOf(e1) && active) rows(B.slot[(k + 1) % LT_BODY_SLOTS], SBK_LT_PFSLOT(par + 1));
        lpfCommit();
        const int b = bodyOf(e0);
        if (b && active) {
            const LBody& bc = B.slot[k % LT_BODY_SLOTS];
            if constexpr (OUT) { SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lFusedOutBody<JT>(c, bc, inst, cy, w, S, sg, vr, vw, SBK_LT_PFSLOT(par)))); }
            else { SBK_DISPATCH_LOCAL(JMASK, bc.joint, (lInwardBody<JT>(c, T, bc, inst, cy, vr, SBK_LT_PFSLOT(par)))); }
        }
        if (e0 & LT_TSYNC) topSync();
        if (e0 & (LT_GSYNC | LT_XSYNC)) groupSync((e0 & LT_XSYNC) != 0);
        ++par; ++k;
        e0 = e1; e1 = e2; e2 = (e1 & LT_END) ? e1 : lst[k + 2];
        lwarpSync();                                               // every lane is done with slot k-1 before it is refilled
    }
}

// Ground's link rows never change: v = 0 in both velocity buffers, a = -g (gravity as a base acceleration).
SBK_HD void lLevelGround(const Ctx& c, const LTables& T, const int inst) {
    constexpr bool BLK = SBK_DEV_BLK;
    SV a0 = zeroSV(); a0.v = mk(-c.gx, -c.gy, -c.gz);
    const CacheRefT<BLK> g = lrecOf<BLK>(c, inst, T.bodies[0].rec);
    g.stSV(lrV(0), zeroSV()); g.stSV(lrV(0) + 6, a0); g.stSV(lrV(0) + LR_VBUF, zeroSV());
}

// One fixed-size RKM step of one warp of the group (lstIn / lstOut: its task lists).  REDUCE(q, u, quat) sums (or, Inf norm,
// maximises) the three error accumulators of this lane's instance over the warps of the group and hands the totals to warp 0,
// which finishes the step for its 32 instances; velValid is group-uniform.
template <int JMASK, class TSYNC, class GSYNC, class REDUCE>
SBK_HD RkmStepResult lListStep(const Ctx& c, const LTables& T, const int* lstIn, const int* lstOut, const LBodySlots& B, const int inst, const bool active,
                               const int lane, double* cy, const LRkmWork& w, const double h, int& vb, const bool velValid, const int wc, int& par,
                               TSYNC topSync, GSYNC groupSync, REDUCE reduce) {
    constexpr bool BLK = SBK_DEV_BLK;
    if (!velValid) lListSweep<JMASK, true>(c, T, lstOut, B, inst, active, lane, cy, w, w.Y, lstageOf(-1, 0.0, w), vb, vb, par, topSync, groupSync);
#pragma unroll 1
    for (int stage = 0; stage < 5; ++stage) {
        const double* S = stage == 0 ? w.Y : w.W;
        const LStage sg = lstageOf(stage, h, w);
        lListSweep<JMASK, false>(c, T, lstIn, B, inst, active, lane, cy, w, S, sg, vb, vb, par, topSync, groupSync);
        // the error sums share carry rows with the inward sweep's P+ hand-over: start them after it
        if (stage == 4 && active) { cy[LF_QACC*SBK_CARRY_STRIDE] = 0; cy[LF_UACC*SBK_CARRY_STRIDE] = 0; cy[LF_QUATACC*SBK_CARRY_STRIDE] = 0; }
        lListSweep<JMASK, true>(c, T, lstOut, B, inst, active, lane, cy, w, S, sg, vb, vb ^ LR_VBUF, par, topSync, groupSync);
        vb ^= LR_VBUF;
    }
    double qAcc = active ? cy[LF_QACC*SBK_CARRY_STRIDE] : 0.0, uAcc = active ? cy[LF_UACC*SBK_CARRY_STRIDE] : 0.0, quatAcc = active ? cy[LF_QUATACC*SBK_CARRY_STRIDE] : 0.0;
    reduce(qAcc, uAcc, quatAcc);
    RkmStepResult res; res.errNorm = 0; res.projected = 0;
    if (active && wc == 0) res = lFinishAttempt<BLK>(c, T, inst, w, qAcc, uAcc, quatAcc);     // one warp finishes its 32 instances
    return res;
}

} // namespace sbkd
