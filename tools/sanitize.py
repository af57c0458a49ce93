#!/usr/bin/env python
"""Reduced run of every integrator / operator kernel family for compute-sanitizer (memcheck, racecheck, synccheck).
Small models and batches: the tools slow kernels down by 10-100x.  Covers: plan 1 body-frame fused integrator with more CTAs
than resident slots (the persistent task queue runs several rounds, release/acquire block flags), its error-controlled
form, the ground-frame integrator (SBK_NOLOCAL), plan 2 (register-resident), plan 3, plan 4 (cooperative grid barrier),
plan 5 (thread-block clusters), and the FULL-record operators."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simbody_b200 as sb
from _harness import ModelInfo

def soa(a): return np.ascontiguousarray(np.asarray(a).T)

def run(name, n, N, plan, steps, h=1e-3, env=None, adaptive=None, ops=False):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k); os.environ[k] = str(v)
    try:
        info = ModelInfo(sb.model_text(name, n))
        q, u = info.random_states(N, 7, q_scale=0.4)
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
        if plan: bm.setPlan(plan)
        bm.setState(soa(q), soa(u), t=0.0)
        if ops:
            bm.realizeAcceleration(); bm.getUDot(); bm.calcMobilizerReactionForces(); bm.calcEnergy()
            bm.multiplyByMInv(np.ones((info.nu, N))); bm.multiplyByM(np.ones((info.nu, N))); bm.calcResidualForceIgnoringConstraints()
        if steps: bm.stepBy(h, steps)
        if adaptive: bm.stepTo(adaptive)
        qq, uu, t = bm.getState(); st, nbad = bm.status()
        assert nbad == 0 and np.all(np.isfinite(qq)), (name, plan)
        print("ok", name, n, N, "plan", bm.getPlan(), "kernel", bm.integratorKernelName(), flush=True)
        bm.close(); topo.close()
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v

big = int(os.environ.get("SBK_SANITIZE_BIG", "40000"))
run("pin_chain", 6, big, 1, 2)                       # > 296 CTAs: several rounds of the task queue
run("humanoid30", 0, 300, 1, 2)
run("humanoid30", 0, 64, 1, 0, adaptive=0.01)
run("welded8", 0, 300, 1, 2)                         # ground-frame integrator (Weld is outside the body-frame set)
run("mixed7", 0, 200, 1, 2, env={"SBK_NOLOCAL": 1})
run("double_pendulum", 0, 1000, 2, 5)
run("mixed7", 0, 40, 3, 2)
run("branched_tree", 60, 64, 4, 1, h=5e-4)
run("twopoint7", 0, 96, 4, 1)
run("branched_tree", 600, 64, 5, 1, h=5e-4)      # two clusters per group, CTA-local levels, cross-cluster barrier
run("mixed7", 0, 100, 1, 0, ops=True)
run("branched_tree", 60, 40, 4, 0, ops=True)
print("SANITIZE_SUBSET_DONE")
