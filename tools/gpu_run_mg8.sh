#!/bin/bash
N=${1:-8}
bash tools/gpu_run_k.sh $N
bash tools/gpu_run_mg.sh $N
