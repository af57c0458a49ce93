#!/bin/bash
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --workload humanoid30_64k --no-extra-workloads 2>/dev/null | grep '^{' > gpurun_out/r2k_h_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2k_h_n$N.json"))
print("n_gpus", d["n_gpus"], "value/gpu %.4g e2e/gpu %.4g" % (d["value"]/d["n_gpus"], d["e2e"]["value"]/d["n_gpus"]), "ms_per_step", d["ms_per_step"], d.get("per_rank"))
PY
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv,noheader
