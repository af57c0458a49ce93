// hostemu.cpp -- TEST-ONLY host build of the kernels' per-body device functions.
//
// The sweeps in simbody_b200/csrc/sbk_sweeps.cuh / sbk_rkm.cuh are __host__ __device__
// templates.  This file compiles them for the CPU and drives them with the same
// thread-per-instance loop structure the CUDA kernels use, so that the kernel *math* can be
// differential-tested against the reference in a container without a GPU.  It is NOT part of
// libsbk.so and is never a fallback: the product library has no CPU path.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <memory>
#include "../../simbody_b200/csrc/topology.h"
#include "../../simbody_b200/csrc/sbk_fused.cuh"
#include "../../simbody_b200/csrc/sbk_lrkm.cuh"
#include "../../simbody_b200/csrc/sbk_ltree.cuh"

using namespace sbkd;

namespace {
struct Emu {
    sbk_topology topo;
    int N = 0;
    long long recTotal = 0;
    std::vector<BodyConst> bodies;
    std::vector<double> cache, y, ydot, qdd, qerr, fmob, Fbody, vin, vout, vin2, Fin;
    std::vector<double> y0, f0, fa, fb, ys;
    std::vector<int> status;
    std::vector<TwoPointConst> tps; std::vector<double> f2;
};
void setup(Emu& e, const char* text, int N) {
    sbk::compileTopology(sbk::fromText(text), e.topo);
    e.N = N; e.bodies = e.topo.bodies;
    long long off = 0;
    for (int b = 0; b < e.topo.nb; ++b) {
        e.bodies[b].cacheBase = off*N;
        off += (b == 0) ? F_H : cacheRecordSize(e.topo.nuOf[b]);
    }
    for (int b = 0; b < e.topo.nb; ++b) e.bodies[b].parentCacheBase = e.bodies[e.bodies[b].parent].cacheBase;
    e.recTotal = off;
    const sbk_topology& t = e.topo;
    e.cache.assign((size_t)off*N, 0.0);
    for (int k = 0; k < N; ++k) { e.cache[(F_XGB+0)*(size_t)N+k] = 1; e.cache[(F_XGB+4)*(size_t)N+k] = 1; e.cache[(F_XGB+8)*(size_t)N+k] = 1; }
    const size_t ny = t.nq + t.nu;
    e.y.assign(ny*N, 0); e.ydot.assign(ny*N, 0); e.qdd.assign((size_t)t.nq*N, 0); e.qerr.assign((size_t)std::max(1, t.nquat)*N, 0);
    e.fmob.assign((size_t)t.nu*N, 0); e.Fbody.assign((size_t)t.nb*6*N, 0);
    e.vin.assign((size_t)t.nu*N, 0); e.vin2.assign((size_t)t.nu*N, 0); e.vout.assign((size_t)t.nu*N, 0); e.Fin.assign((size_t)t.nb*6*N, 0);
    e.y0.assign(ny*N, 0); e.f0.assign(ny*N, 0); e.fa.assign(ny*N, 0); e.fb.assign(ny*N, 0); e.ys.assign(ny*N, 0);
    e.status.assign(N, 0);
    e.tps = t.twoPoint;
    for (TwoPointConst& tp : e.tps) { tp.cacheBase1 = e.bodies[tp.body1].cacheBase; tp.cacheBase2 = e.bodies[tp.body2].cacheBase; }
    e.f2.assign((size_t)t.nb*6*N, 0.0);
}
Ctx makeCtx(Emu& e, int inst) {
    const sbk_topology& t = e.topo; const int N = e.N;
    Ctx c; std::memset(&c, 0, sizeof c);
    c.bodies = e.bodies.data(); c.children = t.children.data(); c.forces = t.forces.data();
    c.nb = t.nb; c.nq = t.nq; c.nu = t.nu; c.nquat = t.nquat;
    c.gx = t.grav[0]; c.gy = t.grav[1]; c.gz = t.grav[2];
    c.cache = e.cache.data(); c.cStride = N; c.cInstStride = 1; c.cSpan = 0; c.cShift = 30; c.cMask = 0x3fffffff;
    c.sStride = N; c.sInstStride = 1; c.sSpan = 0;
    c.q = e.y.data(); c.u = e.y.data() + (size_t)t.nq*N;
    c.qdot = e.ydot.data(); c.udot = e.ydot.data() + (size_t)t.nq*N;
    c.qdotdot = e.qdd.data(); c.qerr = e.qerr.data();
    c.status = e.status.data();
    c.tp = e.tps.data(); c.ntp = (int)e.tps.size(); c.f2 = e.f2.data();
    return c;
}
} // namespace

extern "C" {

int emu_counts(const char* text, int* nb, int* nq, int* nu, int* nquat) {
    try { sbk_topology t; sbk::compileTopology(sbk::fromText(text), t); *nb = t.nb; *nq = t.nq; *nu = t.nu; *nquat = t.nquat; return 0; }
    catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// Same binary layout as `ref_driver eval` (oracle/ref_driver.cpp): instance-major in/out.
int emu_eval(const char* text, int N, const double* in, double* out) {
    try {
        Emu e; setup(e, text, N);
        const sbk_topology& t = e.topo; const int nb = t.nb, nq = t.nq, nu = t.nu, nquat = t.nquat;
        const int inStride = nq + 5*nu + 6*nb;
        const int outStride = nq + nu + nq + nquat + nb*12 + nb*6 + nb*6 + nu + nb*6 + 4*nu + nu + nb*6;
        for (int k = 0; k < N; ++k) {
            const double* p = in + (size_t)k*inStride;
            for (int i = 0; i < nq + nu; ++i) e.y[(size_t)i*N + k] = p[i];
        }
        for (int k = 0; k < N; ++k) {
            const double* p = in + (size_t)k*inStride; double* o = out + (size_t)k*outStride;
            Ctx c = makeCtx(e, k);
            double cy[CARRY_ROWS + LFCARRY_ROWS];
            c.fmobOut = e.fmob.data(); c.FbodyOut = e.Fbody.data();
            tpiEvalDerivatives<false>(c, tablesOf(c), k, cy, c.qdot, c.udot, c.qdotdot);
            if (c.ntp) for (int i = 0; i < 6; ++i) e.Fbody[(size_t)i*N + k] = e.f2[(size_t)i*N + k];      // two-point forces on Ground
            for (int i = 0; i < nq; ++i) *o++ = e.ydot[(size_t)i*N + k];
            for (int i = 0; i < nu; ++i) *o++ = e.ydot[(size_t)(nq+i)*N + k];
            for (int i = 0; i < nq; ++i) *o++ = e.qdd[(size_t)i*N + k];
            for (int i = 0; i < nquat; ++i) *o++ = e.qerr[(size_t)i*N + k];
            auto rec = [&](int b, int f) { return e.cache[(size_t)(e.bodies[b].cacheBase) + (size_t)f*N + k]; };
            for (int b = 0; b < nb; ++b) for (int i = 0; i < 12; ++i) *o++ = rec(b, F_XGB + i);
            for (int b = 0; b < nb; ++b) for (int i = 0; i < 6; ++i)  *o++ = rec(b, F_VGB + i);
            for (int b = 0; b < nb; ++b) for (int i = 0; i < 6; ++i)  *o++ = rec(b, F_AGB + i);
            for (int i = 0; i < nu; ++i) *o++ = e.fmob[(size_t)i*N + k];
            for (int i = 0; i < nb*6; ++i) *o++ = e.Fbody[(size_t)i*N + k];
            // operators
            const double* pa = p + nq + nu; const double* pv = pa + nu; const double* pud = pv + nu;
            const double* pf = pud + nu;    const double* pF = pf + nu;
            c.fmobOut = nullptr; c.FbodyOut = nullptr; c.vecOut = e.vout.data();
            // M*a
            for (int i = 0; i < nu; ++i) e.vin[(size_t)i*N + k] = pa[i];
            c.vecIn = e.vin.data();
            for (int b = 1; b < nb; ++b) idOutDispatch<false>(c, b, k);
            for (int b = nb-1; b >= 1; --b) idInDispatch<false>(c, b, k);
            for (int i = 0; i < nu; ++i) *o++ = e.vout[(size_t)i*N + k];
            // M^-1 v
            for (int i = 0; i < nu; ++i) e.vin[(size_t)i*N + k] = pv[i];
            c.fmobIn = e.vin.data(); c.FbodyIn = nullptr;
            tpiInward<IN_Z>(c, k);
            tpiOutward<false>(c, k, e.vout.data(), nullptr);
            for (int i = 0; i < nu; ++i) *o++ = e.vout[(size_t)i*N + k];
            // residual(f, F, udot)
            for (int i = 0; i < nu; ++i) { e.vin[(size_t)i*N + k] = pud[i]; e.vin2[(size_t)i*N + k] = pf[i]; }
            for (int i = 0; i < nb*6; ++i) e.Fin[(size_t)i*N + k] = pF[i];
            c.vecIn = e.vin.data(); c.fmobIn = e.vin2.data(); c.FbodyIn = e.Fin.data();
            for (int b = 1; b < nb; ++b) idOutDispatch<true>(c, b, k);
            for (int b = nb-1; b >= 1; --b) idInDispatch<true>(c, b, k);
            for (int i = 0; i < nu; ++i) *o++ = e.vout[(size_t)i*N + k];
            // residual with all-zero arguments
            c.vecIn = nullptr; c.fmobIn = nullptr; c.FbodyIn = nullptr;
            for (int b = 1; b < nb; ++b) idOutDispatch<true>(c, b, k);
            for (int b = nb-1; b >= 1; --b) idInDispatch<true>(c, b, k);
            for (int i = 0; i < nu; ++i) *o++ = e.vout[(size_t)i*N + k];
            // calcAcceleration(f, F)
            c.fmobIn = e.vin2.data(); c.FbodyIn = e.Fin.data();
            tpiInward<IN_Z | IN_BIAS>(c, k);
            tpiOutward<true>(c, k, e.vout.data(), nullptr);
            for (int i = 0; i < nu; ++i) *o++ = e.vout[(size_t)i*N + k];
            for (int b = 0; b < nb; ++b) for (int i = 0; i < 6; ++i) *o++ = rec(b, F_AGB + i);
            if (o - (out + (size_t)k*outStride) != outStride) return 3;
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// in: [N][nq+nu]; out: [N][nq+nu+2] (q, u, errNorm of last step, number of projections)
int emu_step(const char* text, int N, const double* in, double* out, double h, int nsteps,
             double accuracy, double consTol, int useInfNorm, int projectEveryStep, int lean) {
    try {
        Emu e; setup(e, text, N);
        const sbk_topology& t = e.topo; const int ny = t.nq + t.nu;
        for (int k = 0; k < N; ++k) for (int i = 0; i < ny; ++i) e.y[(size_t)i*N + k] = in[(size_t)k*ny + i];
        RkmWork w; w.y = e.y.data(); w.y0 = e.y0.data(); w.f0 = e.f0.data(); w.fa = e.fa.data(); w.fb = e.fb.data(); w.ys = e.ys.data();
        w.accuracy = accuracy; w.consTol = consTol; w.useInfNorm = useInfNorm; w.projectEveryStep = projectEveryStep;
        if (lean == 2) {   // register-resident fused plan (2-body Pin/Slider chains only in this emulation)
            if (t.nb != 3) return 4;
            const int j1 = t.bodies[1].joint, j2 = t.bodies[2].joint;
            for (int k = 0; k < N; ++k) {
                double y[8]; for (int i = 0; i < ny; ++i) y[i] = e.y[(size_t)i*N + k];
                double err = 0;
                auto run = [&](auto ch) {
                    ch.b0 = &e.bodies[1]; ch.b1 = &e.bodies[2]; ch.forces = t.forces.data();
                    ch.gx = t.grav[0]; ch.gy = t.grav[1]; ch.gz = t.grav[2];
                    for (int s = 0; s < nsteps; ++s) err = fusedRkmStep(ch, y, h, useInfNorm);
                };
                if (j1 == JT_PIN && j2 == JT_PIN) run(Chain2<JT_PIN, JT_PIN>());
                else if (j1 == JT_SLIDER && j2 == JT_PIN) run(Chain2<JT_SLIDER, JT_PIN>());
                else if (j1 == JT_PIN && j2 == JT_SLIDER) run(Chain2<JT_PIN, JT_SLIDER>());
                else return 4;
                double* o = out + (size_t)k*(ny+2);
                for (int i = 0; i < ny; ++i) o[i] = y[i];
                o[ny] = err; o[ny+1] = 0;
            }
            return 0;
        }
        if (lean >= 3 && !t.localOk) return 5;
        if (lean != 0 && !t.twoPoint.empty()) return 6;      // two-point force elements: FULL records only
        if (lean == 5) {        // the same integrator in level order (sbk_ltree.cuh), one emulated warp group: wc = 0, nw = 1 (or 3: round robin)
            LTables LT; LT.bodies = t.lbodiesLevel.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
            struct { const int* order; const int* start; int nlevels; } LV; LV.order = t.levelOrder.data(); LV.start = t.levelStart.data(); LV.nlevels = t.nlevels;
            for (int k = 0; k < N; ++k) {
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                LRkmWork lw; lw.Y = e.y.data(); lw.W = e.ys.data(); lw.F0 = e.f0.data(); lw.F2 = e.fa.data(); lw.F3 = e.fb.data(); lw.Ynext = e.y.data();
                lw.accuracy = accuracy; lw.consTol = consTol; lw.useInfNorm = useInfNorm; lw.projectEveryStep = projectEveryStep;
                // three emulated warps take the bodies of a level round robin, one after the other (a level's bodies are independent)
                const int NW = 3;
                double cyw[NW][CARRY_ROWS + LFCARRY_ROWS]; int par[NW] = {0, 0, 0};
                lLevelGround(c, LT, k);
                int vb = 0; bool velValid = false; RkmStepResult r; r.errNorm = 0; r.projected = 0; int nproj = 0;
                auto nosync = []() {};
                for (int s = 0; s < nsteps; ++s) {
                    // run the sweeps warp by warp inside each level: emulate by calling the per-level pieces in lockstep
                    auto sweepAll = [&](bool out, const double* S, const LStage& sg, int vr, int vw) {
                        for (int l = out ? 1 : LV.nlevels - 1; out ? l < LV.nlevels : l >= 1; l += out ? 1 : -1) {
                            for (int wcn = 0; wcn < NW; ++wcn) {
                                for (int i = wcn; i < LV.start[l + 1] - LV.start[l]; i += NW) {
                                    const LBody& bc = LT.bodies[LV.order[LV.start[l] + i]];
                                    Ctx cc = c; cc.q = S; cc.u = S + (size_t)cc.nq*cc.sStride;
                                    double* cy = cyw[wcn]; double* pf = cy + LF_PF;
                                    if (out) { lPrefetchOut<JM_MOBILE5, false>(cc, bc, k, pf, lw, S, sg);
                                               SBK_DISPATCH_LOCAL(JM_MOBILE5, bc.joint, (lFusedOutBody<JT>(cc, bc, k, cy, lw, S, sg, vr, vw, pf))); }
                                    else { lPrefetchIn<JM_MOBILE5, false>(cc, bc, k, pf, S, vr);
                                           SBK_DISPATCH_LOCAL(JM_MOBILE5, bc.joint, (lInwardBody<JT>(cc, LT, bc, k, cy, vr, pf))); }
                                }
                            }
                        }
                    };
                    if (!velValid) sweepAll(true, lw.Y, lstageOf(-1, 0.0, lw), vb, vb);
                    for (int stage = 0; stage < 5; ++stage) {
                        const double* S = stage == 0 ? lw.Y : lw.W;
                        if (stage == 4) for (int wcn = 0; wcn < NW; ++wcn) { cyw[wcn][LF_QACC] = 0; cyw[wcn][LF_UACC] = 0; cyw[wcn][LF_QUATACC] = 0; }
                        sweepAll(false, S, lstageOf(stage, h, lw), vb, vb);
                        sweepAll(true, S, lstageOf(stage, h, lw), vb, vb ^ LR_VBUF);
                        vb ^= LR_VBUF;
                    }
                    double qa = 0, ua = 0, qt = 0;
                    for (int wcn = 0; wcn < NW; ++wcn) {
                        if (useInfNorm) { qa = normMax(qa, cyw[wcn][LF_QACC]); ua = normMax(ua, cyw[wcn][LF_UACC]); qt = normMax(qt, cyw[wcn][LF_QUATACC]); }
                        else { qa += cyw[wcn][LF_QACC]; ua += cyw[wcn][LF_UACC]; qt += cyw[wcn][LF_QUATACC]; }
                    }
                    r = lFinishAttempt<false>(c, LT, k, lw, qa, ua, qt);
                    velValid = !r.projected; nproj += r.projected;
                }
                (void)nosync; (void)par;
                double* o = out + (size_t)k*(ny+2);
                for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
                o[ny] = r.errNorm; o[ny+1] = nproj;
            }
            return 0;
        }
        if (lean == 6) {        // lListStep itself with one emulated warp: its task lists hold the subtree walks below the cut, then the top levels
            const sbk::TreeCut cut = sbk::cutTreeForWarps(t, 1, 1);
            const sbk::TreeCut cut4 = sbk::cutTreeForWarps(t, 4, 1);     // flags / walks of a cut at the first level >= 4 wide, still one warp
            // one warp executing the schedule of a 4-warp cut: concatenate the four warps' subtree lists, then the top
            std::vector<int> lin, lout;
            { const sbk::TreeCut& q = cut4; const int nw = 4;
              auto seg = [&](int dir, int w) { std::vector<int> v; for (int k = q.listStart[(size_t)dir*nw + w]; !(q.lists[k] & LT_END); ++k) v.push_back(q.lists[k]); return v; };
              // inward: all subtree parts (entries before the GSYNC flag, inclusive) of every warp, then the top part of warp 0
              std::vector<int> top_in, top_out;
              for (int w = 0; w < nw; ++w) { std::vector<int> v = seg(0, w); size_t i = 0; bool split = false;
                  for (; i < v.size(); ++i) { const bool gs = (v[i] & LT_GSYNC) != 0; lin.push_back(v[i] & ~(LT_GSYNC | LT_TSYNC)); if (gs) { split = true; ++i; break; } }
                  if (w == 0) for (; i < v.size(); ++i) top_in.push_back(v[i] & ~(LT_GSYNC | LT_TSYNC));
                  (void)split; }
              lin.insert(lin.end(), top_in.begin(), top_in.end());
              for (int w = 0; w < nw; ++w) { std::vector<int> v = seg(1, w); size_t i = 0;
                  if (w == 0) { for (; i < v.size(); ++i) { const bool gs = (v[i] & LT_GSYNC) != 0; top_out.push_back(v[i] & ~(LT_GSYNC | LT_TSYNC)); if (gs) { ++i; break; } } }
                  else { for (; i < v.size(); ++i) if (v[i] & LT_GSYNC) { ++i; break; } } 
                  std::vector<int> rest(v.begin() + i, v.end());
                  if (w == 0) lout.insert(lout.begin(), top_out.begin(), top_out.end());
                  for (int x : rest) lout.push_back(x & ~(LT_GSYNC | LT_TSYNC)); }
              for (int k = 0; k < 3; ++k) { lin.push_back(LT_END); lout.push_back(LT_END); } }
            (void)cut;
            LTables LT; LT.bodies = cut4.bodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
            for (int k = 0; k < N; ++k) {
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                LRkmWork lw; lw.Y = e.y.data(); lw.W = e.ys.data(); lw.F0 = e.f0.data(); lw.F2 = e.fa.data(); lw.F3 = e.fb.data(); lw.Ynext = e.y.data();
                lw.accuracy = accuracy; lw.consTol = consTol; lw.useInfNorm = useInfNorm; lw.projectEveryStep = projectEveryStep;
                double cy[CARRY_ROWS + LFCARRY_ROWS]; int par = 0, vb = 0; bool velValid = false; int nproj = 0;
                LBody slots[LT_BODY_SLOTS]; LBodySlots BS; BS.slot = slots;
                lLevelGround(c, LT, k);
                RkmStepResult r; r.errNorm = 0; r.projected = 0;
                for (int s = 0; s < nsteps; ++s) {
                    r = lListStep<JM_MOBILE5>(c, LT, lin.data(), lout.data(), BS, k, true, 0, cy, lw, h, vb, velValid, 0, par, []() {}, [](bool) {}, [](double&, double&, double&) {});
                    velValid = !r.projected; nproj += r.projected;
                }
                double* o = out + (size_t)k*(ny+2);
                for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
                o[ny] = r.errNorm; o[ny+1] = nproj;
            }
            return 0;
        }
        if (lean == 4) {        // fused two-sweep integrator on the body-frame cores (sbk_lrkm.cuh); Ynext == Y (fixed step)
            LTables LT; LT.bodies = t.lbodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
            for (int k = 0; k < N; ++k) {
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                double cy[CARRY_ROWS + LFCARRY_ROWS];
                LRkmWork lw; lw.Y = e.y.data(); lw.W = e.ys.data(); lw.F0 = e.f0.data(); lw.F2 = e.fa.data(); lw.F3 = e.fb.data(); lw.Ynext = e.y.data();
                lw.accuracy = accuracy; lw.consTol = consTol; lw.useInfNorm = useInfNorm; lw.projectEveryStep = projectEveryStep;
                LRkmState st; st.vb = 0; st.velValid = false;
                RkmStepResult r; r.errNorm = 0; r.projected = 0; int nproj = 0;
                for (int s = 0; s < nsteps; ++s) { r = lRkmAttempt<JM_MOBILE5>(c, LT, k, lw, h, cy, st); nproj += r.projected; }
                double* o = out + (size_t)k*(ny+2);
                for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
                o[ny] = r.errNorm; o[ny+1] = nproj;
            }
            return 0;
        }
        LTables LT; LT.bodies = t.lbodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
        for (int k = 0; k < N; ++k) {
            Ctx c = makeCtx(e, k);
            c.qdotdot = nullptr; c.qerr = nullptr;
            double cy[CARRY_ROWS + LFCARRY_ROWS];
            RkmStepResult r; r.errNorm = 0; r.projected = 0; int nproj = 0;
            for (int s = 0; s < nsteps; ++s) {
                r = lean == 3 ? tpiRkmStep<true, JM_ALL>(c, LT, k, w, h, cy)          // body-frame sweeps (sbk_local.cuh)
                  : lean ? tpiRkmStep<true>(c, tablesOf(c), k, w, h, cy) : tpiRkmStep<false>(c, tablesOf(c), k, w, h, cy);
                nproj += r.projected;
            }
            double* o = out + (size_t)k*(ny+2);
            for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
            o[ny] = r.errNorm; o[ny+1] = nproj;
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// Derivatives only: in [N][nq+nu] -> out [N][nq+nu] (qdot, udot).  local = 1: body-frame sweeps, 0: LEAN ground-frame sweeps.
int emu_deriv(const char* text, int N, const double* in, double* out, int local) {
    try {
        Emu e; setup(e, text, N);
        const sbk_topology& t = e.topo; const int ny = t.nq + t.nu;
        if (local && !t.localOk) return 5;
        for (int k = 0; k < N; ++k) for (int i = 0; i < ny; ++i) e.y[(size_t)i*N + k] = in[(size_t)k*ny + i];
        LTables LT; LT.bodies = t.lbodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
        for (int k = 0; k < N; ++k) {
            Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
            double cy[CARRY_ROWS + LFCARRY_ROWS];
            if (local) lEvalDerivatives(c, LT, k, cy, c.qdot, c.udot);
            else tpiEvalDerivatives<true>(c, tablesOf(c), k, cy, c.qdot, c.udot, nullptr);
            for (int i = 0; i < ny; ++i) out[(size_t)k*ny + i] = e.ydot[(size_t)i*N + k];
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// Error-controlled stepping; out [N][ny+5]: y | steps | attempts | steps+4*attempts | last step | advanced time
int emu_adaptive(const char* text, int N, const double* in, double* out, double tFinal, double accuracy, double initStep,
                 int allowInterpolation, int fused) {
    try {
        Emu e; setup(e, text, N);
        const sbk_topology& t = e.topo; const int ny = t.nq + t.nu;
        for (int k = 0; k < N; ++k) for (int i = 0; i < ny; ++i) e.y[(size_t)i*N + k] = in[(size_t)k*ny + i];
        RkmWork w; w.y = e.y.data(); w.y0 = e.y0.data(); w.f0 = e.f0.data(); w.fa = e.fa.data(); w.fb = e.fb.data(); w.ys = e.ys.data();
        w.accuracy = accuracy; w.consTol = accuracy/10; w.useInfNorm = 0; w.projectEveryStep = 0;
        StepLimits lim; lim.accuracy = accuracy; lim.minStep = -1; lim.maxStep = -1;
        for (int k = 0; k < N; ++k) {
            AdaptiveState st; st.t = 0; st.h = initStep; st.lastStep = initStep; st.steps = 0; st.attempts = 0;
            double lastErr = 0; int nproj = 0;
            double* o = out + (size_t)k*(ny+5);
            if (fused == 1) {
                if (t.nb != 3 || t.bodies[1].joint != JT_PIN || t.bodies[2].joint != JT_PIN) return 4;
                double y[8]; for (int i = 0; i < ny; ++i) y[i] = e.y[(size_t)i*N + k];
                Chain2<JT_PIN, JT_PIN> ch; ch.b0 = &e.bodies[1]; ch.b1 = &e.bodies[2]; ch.forces = t.forces.data();
                ch.gx = t.grav[0]; ch.gy = t.grav[1]; ch.gz = t.grav[2];
                fusedRkmAdaptive(ch, y, lim, tFinal, allowInterpolation, 1000000, 0, st, lastErr);
                for (int i = 0; i < ny; ++i) o[i] = y[i];
            } else if (fused == 2 || fused == 3) {   // fused two-sweep integrator on the body-frame cores; 3: the CTA-voting lockstep form
                if (!t.localOk) return 5;
                LTables LT; LT.bodies = t.lbodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                double cy[CARRY_ROWS + LFCARRY_ROWS];
                LRkmWork lw; lw.Y = e.y.data(); lw.W = e.ys.data(); lw.F0 = e.f0.data(); lw.F2 = e.fa.data(); lw.F3 = e.fb.data(); lw.Ynext = e.y0.data();
                lw.accuracy = accuracy; lw.consTol = accuracy/10; lw.useInfNorm = 0; lw.projectEveryStep = 0;
                LRkmState ls; ls.vb = 0; ls.velValid = false;
                if (fused == 3) lRkmAdaptive<JM_MOBILE5, CtaVote>(c, LT, k, lw, lim, tFinal, allowInterpolation, 1000000, st, cy, ls, lastErr, nproj, true);
                else lRkmAdaptive<JM_MOBILE5>(c, LT, k, lw, lim, tFinal, allowInterpolation, 1000000, st, cy, ls, lastErr, nproj);
                for (int i = 0; i < ny; ++i) o[i] = lw.Y[(size_t)i*N + k];
            } else {
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                double cy[CARRY_ROWS + LFCARRY_ROWS];
                if (fused == 4) tpiRkmAdaptive<true, JM_ALL, Tables, CtaVote>(c, tablesOf(c), k, w, lim, tFinal, allowInterpolation, 1000000, st, cy, lastErr, nproj, true);
                else tpiRkmAdaptive<true>(c, tablesOf(c), k, w, lim, tFinal, allowInterpolation, 1000000, st, cy, lastErr, nproj);
                for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
            }
            o[ny] = st.steps; o[ny+1] = st.attempts; o[ny+2] = st.steps + 4.0*st.attempts; o[ny+3] = st.lastStep; o[ny+4] = st.t;
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

int emu_model_text(const char* name, int n, char* buf, int cap) {
    try {
        const std::string s = sbk::toText(sbk::makeNamedModel(name, n));
        if ((int)s.size() + 1 <= cap) std::memcpy(buf, s.c_str(), s.size() + 1);
        return (int)s.size() + 1;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return -1; }
}

// Schedule check of the plan-5 task lists (topology.cpp: cutTreeForWarps) by simulation: warps run their lists, a barrier entry
// holds a warp until every warp of its scope (CTA of 8 warps / cluster / all clusters) waits at a barrier of the same kind.
// Returns 0 if every body is processed exactly once per direction, a child always strictly before (inward) / after (outward)
// its parent with a barrier or the same warp in between, and no warp is left waiting; else a positive error code.
int emu_cut_check(const char* text, int nwarps, int topWarps, int cutWidth, int nclusters) {
    try {
        sbk_topology t; sbk::compileTopology(sbk::fromText(text), t);
        const sbk::TreeCut cut = sbk::cutTreeForWarps(t, nwarps, topWarps, cutWidth, nclusters);
        const int perCluster = nwarps/std::max(1, nclusters);
        for (int dir = 0; dir < 2; ++dir) {
            std::vector<int> pos(nwarps), doneRound(t.nb, -1), doneWarp(t.nb, -1), waiting(nwarps, 0);
            for (int w = 0; w < nwarps; ++w) pos[w] = cut.listStart[(size_t)dir*nwarps + w];
            std::vector<char> finished(nwarps, 0);
            doneRound[0] = -2; doneWarp[0] = -2;       // Ground: always available
            for (int round = 0; ; ++round) {
                bool progress = false;
                for (int w = 0; w < nwarps; ++w) {
                    if (finished[w] || waiting[w]) continue;
                    for (;;) {
                        const int e = cut.lists[pos[w]];
                        if (e & sbkd::LT_END) { finished[w] = 1; progress = true; break; }
                        ++pos[w]; progress = true;
                        const int b = e & sbkd::LT_BODY_MASK;
                        if (b) {
                            if (b >= t.nb || doneRound[b] != -1) return 2;                    // unknown body / processed twice
                            const sbkd::LBody& lb = t.lbodies[b];
                            if (dir == 0) {
                                for (int j = 0; j < lb.nchild; ++j) {
                                    const int ch = t.children[lb.childStart + j];
                                    if (doneRound[ch] == -1) return 3;                        // child not done yet
                                    if (doneRound[ch] == round && doneWarp[ch] != w) return 4; // concurrent with another warp's child
                                }
                            } else {
                                const int pa = lb.parent;
                                if (doneRound[pa] == -1) return 3;
                                if (doneRound[pa] == round && doneWarp[pa] != w) return 4;
                            }
                            doneRound[b] = round; doneWarp[b] = w;
                        }
                        const int f = e & (sbkd::LT_TSYNC | sbkd::LT_GSYNC | sbkd::LT_XSYNC);
                        if (f) { waiting[w] = (f & sbkd::LT_XSYNC) ? 3 : (f & sbkd::LT_GSYNC) ? 2 : 1; break; }
                    }
                }
                // release the barriers whose whole scope has arrived
                auto release = [&](int lo, int hi, int kind) {
                    for (int w = lo; w < hi; ++w) if (waiting[w] != kind) return;
                    for (int w = lo; w < hi; ++w) waiting[w] = 0;
                    progress = true;
                };
                for (int w0 = 0; w0 < nwarps; w0 += 8) release(w0, std::min(nwarps, w0 + 8), 1);
                for (int w0 = 0; w0 < nwarps; w0 += perCluster) release(w0, std::min(nwarps, w0 + perCluster), 2);
                release(0, nwarps, 3);
                bool all = true; for (int w = 0; w < nwarps; ++w) all = all && finished[w];
                if (all) break;
                if (!progress) return 5;                                                       // a warp waits for ever
            }
            for (int b = 1; b < t.nb; ++b) if (doneRound[b] == -1) return 6;                   // body never processed
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// Shape of a plan-5 schedule: out = {cut level, entries / TSYNC / GSYNC / XSYNC counts of the longest inward list}.
int emu_cut_info(const char* text, int nwarps, int topWarps, int cutWidth, int nclusters, int* out) {
    try {
        sbk_topology t; sbk::compileTopology(sbk::fromText(text), t);
        const sbk::TreeCut cut = sbk::cutTreeForWarps(t, nwarps, topWarps, cutWidth, nclusters);
        out[0] = cut.cutLevel; out[1] = out[2] = out[3] = out[4] = 0;
        for (int w = 0; w < nwarps; ++w) {
            int n = 0, ts = 0, gs = 0, xs = 0;
            for (int k = cut.listStart[w]; !(cut.lists[k] & sbkd::LT_END); ++k) {
                ++n; ts += (cut.lists[k] & sbkd::LT_TSYNC) != 0; gs += (cut.lists[k] & sbkd::LT_GSYNC) != 0; xs += (cut.lists[k] & sbkd::LT_XSYNC) != 0;
            }
            if (n > out[1]) { out[1] = n; out[2] = ts; out[3] = gs; out[4] = xs; }
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

// Plan 5 with its real schedule on the host: one std::thread per warp of the group runs lListStep (sbk_ltree.cuh) over the
// task lists of cutTreeForWarps, with cyclic barriers standing in for __syncthreads (8 warps), barrier.cluster and the
// cross-cluster barrier, and the fixed-order reduction of the error sums -- what ctreeRkmKernel does for one lane.
namespace {
struct CyclicBarrier {
    std::mutex m; std::condition_variable cv; int n, waiting = 0; long gen = 0;
    explicit CyclicBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
}
int emu_cluster_step(const char* text, int N, const double* in, double* out, double h, int nsteps, double accuracy,
                     int nwarps, int topWarps, int cutWidth, int nclusters) {
    try {
        Emu e; setup(e, text, N);
        const sbk_topology& t = e.topo; const int ny = t.nq + t.nu;
        if (!t.localOk) return 5;
        for (int k = 0; k < N; ++k) for (int i = 0; i < ny; ++i) e.y[(size_t)i*N + k] = in[(size_t)k*ny + i];
        const sbk::TreeCut cut = sbk::cutTreeForWarps(t, nwarps, topWarps, cutWidth, nclusters);
        LTables LT; LT.bodies = cut.bodies.data(); LT.children = t.children.data(); LT.forces = t.forces.data(); LT.fcoef = t.lfcoef.data();
        const int perCluster = nwarps/std::max(1, nclusters);
        for (int k = 0; k < N; ++k) {
            std::vector<std::unique_ptr<CyclicBarrier>> ctaBar, clusterBar;
            for (int w0 = 0; w0 < nwarps; w0 += 8) ctaBar.emplace_back(new CyclicBarrier(std::min(8, nwarps - w0)));
            for (int w0 = 0; w0 < nwarps; w0 += perCluster) clusterBar.emplace_back(new CyclicBarrier(std::min(perCluster, nwarps - w0)));
            CyclicBarrier allBar(nwarps);
            std::vector<double> part((size_t)nwarps*3, 0.0);
            int projectedFlag = 0; RkmStepResult last; last.errNorm = 0; last.projected = 0; int nprojTotal = 0;
            { Ctx c0 = makeCtx(e, k); lLevelGround(c0, LT, k); }
            auto warp = [&](const int wc) {
                Ctx c = makeCtx(e, k); c.qdotdot = nullptr; c.qerr = nullptr;
                LRkmWork lw; lw.Y = e.y.data(); lw.W = e.ys.data(); lw.F0 = e.f0.data(); lw.F2 = e.fa.data(); lw.F3 = e.fb.data(); lw.Ynext = e.y.data();
                lw.accuracy = accuracy; lw.consTol = accuracy/10; lw.useInfNorm = 0; lw.projectEveryStep = 0;
                std::vector<double> cyv(CARRY_ROWS + LFCARRY_ROWS, 0.0); double* cy = cyv.data();
                LBody slots[LT_BODY_SLOTS]; LBodySlots BS; BS.slot = slots;
                const int* lstIn = cut.lists.data() + cut.listStart[wc]; const int* lstOut = cut.lists.data() + cut.listStart[(size_t)nwarps + wc];
                int par = 0, vb = 0; bool velValid = false;
                auto topSync = [&]() { ctaBar[wc/8]->wait(); };
                auto groupSync = [&](const bool cross) { if (cross && nclusters > 1) allBar.wait(); else clusterBar[wc/perCluster]->wait(); };
                auto reduce = [&](double& q, double& u, double& qt) {
                    part[(size_t)wc*3] = q; part[(size_t)wc*3 + 1] = u; part[(size_t)wc*3 + 2] = qt;
                    allBar.wait();
                    if (wc == 0) { double s0 = 0, s1 = 0, s2 = 0; for (int w = 0; w < nwarps; ++w) { s0 += part[(size_t)w*3]; s1 += part[(size_t)w*3 + 1]; s2 += part[(size_t)w*3 + 2]; } q = s0; u = s1; qt = s2; }
                };
                for (int s = 0; s < nsteps; ++s) {
                    const RkmStepResult r = lListStep<JM_MOBILE5>(c, LT, lstIn, lstOut, BS, k, true, 0, cy, lw, h, vb, velValid, wc, par, topSync, groupSync, reduce);
                    if (wc == 0) { projectedFlag = r.projected; last = r; nprojTotal += r.projected; }
                    allBar.wait();
                    velValid = projectedFlag == 0;
                    allBar.wait();          // nobody overwrites the flag before everyone has read it
                }
            };
            std::vector<std::thread> th;
            for (int w = 0; w < nwarps; ++w) th.emplace_back(warp, w);
            for (auto& x : th) x.join();
            double* o = out + (size_t)k*(ny+2);
            for (int i = 0; i < ny; ++i) o[i] = e.y[(size_t)i*N + k];
            o[ny] = last.errNorm; o[ny+1] = nprojTotal;
        }
        return 0;
    } catch (const std::exception& ex) { std::fprintf(stderr, "emu: %s\n", ex.what()); return 1; }
}

} // extern "C"
