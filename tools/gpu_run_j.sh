#!/bin/bash
mkdir -p gpurun_out
python tools/quick_perf.py pin_chain50_64k humanoid30_64k branched_tree1000_256 2>&1 | tee gpurun_out/r2j_perf.txt
SBK_SPL=37 python tools/quick_perf.py pin_chain50_64k humanoid30_64k 2>&1 | tee -a gpurun_out/r2j_perf.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
