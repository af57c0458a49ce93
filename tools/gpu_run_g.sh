#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cluster_level or tree_1000 or auto_plan or projection_options or status_word or adaptive_runs" > gpurun_out/r2g_pytest.txt 2>&1
tail -15 gpurun_out/r2g_pytest.txt
for cs in 8 4; do SBK_CLUSTER=$cs SBK_TAG=cluster$cs timeout 120 python tools/quick_perf.py branched_tree1000_256 2>&1 | tail -2; done | tee gpurun_out/r2g_perf.txt
