#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1
tail -30 gpurun_out/r2e_pytest.txt
python tools/quick_perf.py pin_chain50_64k humanoid30_64k 2>&1 | tee gpurun_out/r2e_perf.txt
