#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_run_san.sh
