#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests -x -q -m gpu 2>&1 | tail -3
  SBK_SPL=111 python tools/quick_perf.py pin_chain50_64k
  SBK_SPL=37 python tools/quick_perf.py humanoid30_64k
) 2>&1 | tee gpurun_out/r2k_suite.log
