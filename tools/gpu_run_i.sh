#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 64" "64 64" "64 128" "64 256" "16 64" "32 128"; do
  set -- $cfg
  SBK_TOPWARPS=$1 SBK_CUTWARPS=$2 SBK_TAG="top$1_cut$2" timeout 120 python tools/quick_perf.py branched_tree1000_256 2>&1 | tail -1
done | tee gpurun_out/r2i_perf.txt
SBK_TOPWARPS=64 SBK_CUTWARPS=128 timeout 600 python -m pytest tests -m gpu -x -q -k "cluster_level or tree_1000 or status_word" 2>&1 | tail -3
