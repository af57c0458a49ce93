#!/bin/bash
# multi-GPU checks: torchrun + NCCL (timing barrier / max, end-of-run statistics, all-gather of the final states), weak and strong scaling
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_weak_n$N.json 2> gpurun_out/r2_scale_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-extra-workloads > gpurun_out/r2_scale_strong_c2_n$N.json 2>> gpurun_out/r2_scale_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --workload humanoid30_64k --no-extra-workloads > gpurun_out/r2_scale_weak_humanoid_n$N.json 2>> gpurun_out/r2_scale_n$N.err
tail -5 gpurun_out/r2_scale_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2_scale_weak_n$N.json", "gpurun_out/r2_scale_strong_c2_n$N.json", "gpurun_out/r2_scale_weak_humanoid_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, "n_gpus", d["n_gpus"], "scaling", d["scaling"], "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), "inst/gpu", d["config"]["instances_per_gpu"], "gather", d.get("final_states_all_gather"))
        for k, v in d.get("workloads", {}).items(): print("   ", k, "value %.4g e2e %.4g frac %.3f" % (v["value"], v["e2e"], v["fp64_frac"]))
    except Exception as e:
        print(f, "ERR", e)
PY
