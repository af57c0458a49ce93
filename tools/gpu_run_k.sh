#!/bin/bash
python - <<'PY'
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, simbody_b200 as sb
from _harness import ModelInfo
from bench import WORKLOADS
for name, env in (("pin_chain50_64k", {}), ("humanoid30_64k", {})):
    wl = WORKLOADS[name]; info = ModelInfo(sb.model_text(wl["model"], wl["n"])); N = wl["batch"]
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
    q, u = info.random_states(N, 12345, q_scale=wl["q_scale"])
    bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
    ms = []
    for _ in range(8):
        bm.stepBy(wl["h"], wl["spl"]); ms.append(round(bm.lastKernelMs(), 2))
    print(name, env, "mean %.2f" % np.mean(ms[1:]), ms, flush=True)
    bm.close(); topo.close()
PY
python -m pytest tests -x -q -m gpu -k "pin or chain or full_size" 2>&1 | tail -3
