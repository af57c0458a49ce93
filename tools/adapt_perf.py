#!/usr/bin/env python
"""Throughput of the error-controlled integrator (tuning aid). usage: tools/adapt_perf.py workload [t_final] [batch]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simbody_b200 as sb
from _harness import ModelInfo
from bench import WORKLOADS
name = sys.argv[1]; wl = WORKLOADS[name]
tf = float(sys.argv[2]) if len(sys.argv) > 2 else wl.get("t_adapt", 0.05)
N = int(sys.argv[3]) if len(sys.argv) > 3 else wl["batch"]
info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
q, u = info.random_states(N, 12345, q_scale=wl["q_scale"])
for rep in range(2):
    bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
    st, at, last = bm.stepTo(tf, accuracy=1e-3, init_step=wl["h"])
    ms = bm.lastKernelMs()
print(json.dumps({"workload": name, "N": N, "t_final": tf, "ms": ms, "accepted_per_s": float(st.sum())/(ms*1e-3), "mean": float(st.mean()), "max": int(st.max()),
                  "attempts_mean": float(at.mean()), "warp_max_mean": float(st.reshape(-1, 32).max(axis=1).mean()), "nolocal": os.environ.get("SBK_NOLOCAL", ""),
                  "qproj": bm.stats()["q_projections"]}))
