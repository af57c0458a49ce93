#!/bin/bash
python tools/adapt_perf.py humanoid30_64k 0.05
SBK_NOLOCAL=1 python tools/adapt_perf.py humanoid30_64k 0.05
python tools/adapt_perf.py pin_chain50_64k 0.05
python tools/adapt_perf.py double_pendulum_1M 0.5
python -m pytest tests -x -q -m gpu -k "adaptive" 2>&1 | tail -3
