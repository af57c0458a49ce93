#!/usr/bin/env python
"""Quick device-resident throughput of the fixed-step integrator for tuning (not the bench contract).
usage: tools/quick_perf.py [workload ...]  ; env SBK_NOLOCAL=1 selects the ground-frame integrator."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simbody_b200 as sb
from _harness import ModelInfo
from bench import WORKLOADS, algorithmic_work

def run(name, spl, reps=3, plan=None, batch=None):
    wl = WORKLOADS[name]
    info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
    N = batch or wl["batch"]
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
    if plan: bm.setPlan(plan)
    q, u = info.random_states(N, 12345, q_scale=wl["q_scale"])
    bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
    bm.stepBy(wl["h"], spl); bm.lib.sbk_synchronize(bm.handle)
    best = 1e30
    for _ in range(reps):
        bm.stepBy(wl["h"], spl); ms = bm.lastKernelMs(); best = min(best, ms)
    flop, _ = algorithmic_work(info)
    rate = N*spl/(best*1e-3)
    st, nbad = bm.status()
    print(json.dumps({"workload": name, "N": N, "plan": bm.getPlan(), "nolocal": os.environ.get("SBK_NOLOCAL", ""), "ms": round(best, 3), "inst_steps_per_s": rate,
                      "alg_tflops": rate*flop/1e12, "nbad": int(nbad), "tag": os.environ.get("SBK_TAG", "")}), flush=True)
    bm.close(); topo.close()

if __name__ == "__main__":
    names = sys.argv[1:] or ["pin_chain50_64k", "humanoid30_64k"]
    for n in names:
        spl = int(os.environ.get("SBK_SPL", 0)) or {"double_pendulum_1M": 100, "pin_chain50_64k": 8, "humanoid30_64k": 8, "branched_tree1000_256": 8}[n]
        run(n, spl)
