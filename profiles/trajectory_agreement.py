#!/usr/bin/env python
"""Trajectory agreement GPU engine vs the reference (oracle/_ref), fixed-step RKM, as a function of
the horizon.  Run on a GPU box:  python profiles/trajectory_agreement.py > gpurun_out/trajectory.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simbody_b200 as sb
from _harness import ModelInfo, RefDriver

out = {}
for name, n, h, horizons, qs in [("double_pendulum", 0, 1e-3, [0.1, 0.5, 1, 2, 5, 10, 20], 2.0),
                                 ("pin_chain", 50, 1e-3, [0.05, 0.2, 0.5], 1.0),
                                 ("humanoid30", 0, 1e-3, [0.05, 0.2, 0.5, 1.0], 0.4),
                                 ("branched_tree", 1000, 5e-4, [0.005, 0.02, 0.05], 0.4)]:
    info = ModelInfo(sb.model_text(name, n))
    N = 32
    q, u = info.random_states(N, 777, q_scale=qs)
    y0 = np.concatenate([q, u], axis=1)
    ny = info.nq + info.nu
    ref = RefDriver()
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, N)
    bm.setStateAoS(q, u)
    done = 0; rows = []
    for T in horizons:
        steps = int(round(T / h))
        bm.stepBy(h, steps - done); done = steps
        qg, ug = bm.getStateAoS()
        yr = ref.step(info, y0, h, steps)[:, :ny]
        yg = np.concatenate([qg, ug], axis=1)
        err = np.max(np.abs(yg - yr), axis=1) / np.maximum(1.0, np.max(np.abs(yr), axis=1))
        rows.append({"t": T, "steps": steps, "max_rel_err": float(err.max()), "median_rel_err": float(np.median(err))})
    out[name] = {"h": h, "instances": N, "plan": bm.getPlan(), "agreement": rows}
    bm.close(); topo.close()
print(json.dumps(out, indent=1))
