#!/bin/bash
mkdir -p gpurun_out
python tools/quick_perf.py pin_chain50_64k humanoid30_64k branched_tree1000_256 double_pendulum_1M 2>&1 | tee gpurun_out/r2d_perf.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
