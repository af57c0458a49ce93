#!/usr/bin/env python
"""Error-controlled kernel forced to a constant step (min_step = max_step = h): per-step cost against the fixed-step kernel."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simbody_b200 as sb
from _harness import ModelInfo
from bench import WORKLOADS
name = sys.argv[1]; wl = WORKLOADS[name]; nst = int(sys.argv[2]) if len(sys.argv) > 2 else 37
N = wl["batch"]; h = wl["h"]
info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
q, u = info.random_states(N, 12345, q_scale=wl["q_scale"])
for rep in range(2):
    bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
    st, at, last = bm.stepTo(nst*h, accuracy=1e-3, init_step=h, min_step=h, max_step=h)
    ms = bm.lastKernelMs()
bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
bm.stepBy(h, nst); bm.stepBy(h, nst); msf = bm.lastKernelMs()
print(json.dumps({"workload": name, "steps": nst, "adaptive_ms": ms, "fixed_ms": msf, "ratio": ms/msf, "mean": float(st.mean()), "max": int(st.max()), "att": float(at.mean())}))
