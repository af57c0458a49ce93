#!/usr/bin/env python
"""bench.py -- batched RK-Merson instance-steps/s on B200 (BASELINE.json metric).

A "step" of this benchmark is ONE launch of the hot path over the batch: `--steps-per-launch`
fixed-size Runge-Kutta-Merson steps (5 derivative evaluations each) for every instance.
  value : whole-job instance-steps/s, state resident in HBM, CUDA events on the launch stream
  e2e   : the same through the public C ABI with HOST buffers: sbk_set_state (H2D) +
          sbk_rkm_step + sbk_get_state (D2H) inside the timed region
  --impl reference : the reference's own CPU path (oracle/_ref, real Simbody, one System per
          host thread) on the same workload, bounded sample.
One process per GPU under torchrun; instances shard across ranks with no data-path collective
(weak scaling: per-GPU batch fixed).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# workload name -> (model, size param, per-GPU batch, h, algorithmic flop / instance-step, q scale)
# flop/instance-step = 5 * sum_bodies F_eval(joint) + 30*ny  
WORKLOADS = {
    "double_pendulum_1M": dict(model="double_pendulum", n=0, batch=1048576, h=1e-3, q_scale=3.0),
    "pin_chain50_64k":    dict(model="pin_chain", n=50, batch=65536, h=1e-3, q_scale=1.0),
    "humanoid30_64k":     dict(model="humanoid30", n=0, batch=65536, h=1e-3, q_scale=0.5),
    "branched_tree1000_256": dict(model="branched_tree", n=1000, batch=256, h=5e-4, q_scale=0.5),
}
# measured DRAM bytes per instance-step (ncu --set full, profiles/r1_prof_<workload>.txt): launch traffic / (N * steps)
NCU_TRAFFIC_PER_INSTANCE_STEP = {"double_pendulum_1M": 4.276e7 / (1048576 * 20), "humanoid30_64k": 2.726e10 / (65536 * 4),
                                 "pin_chain50_64k": 1.407e10 / (65536 * 4),
                                 "branched_tree1000_256": 1.762e9 / 256}
# sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active of the same captures (the hardware's own
# view of FP64-pipe utilisation; roofline.frac below uses the reference's ALGORITHMIC flop count instead)
NCU_FP64_PIPE_ACTIVE = {"double_pendulum_1M": 0.776, "humanoid30_64k": 0.246, "pin_chain50_64k": 0.335, "branched_tree1000_256": 0.037}
F_EVAL = {"PIN": 1180.0, "SLIDER": 1130.0, "UNIVERSAL": 1710.0, "BALL": 2060.0, "FREE": 3800.0, "WELD": 700.0,
          "TRANSLATION": 1800.0, "CYLINDER": 1650.0, "PLANAR": 1950.0, "GIMBAL": 2300.0}


def algorithmic_work(info):
    """(flop, compulsory HBM bytes) per instance-step, SURVEY.md section 8(d): the reference's own flop
    accounting per derivative evaluation and mobilizer kind for general frames; a body whose R_PF is
    exactly the identity (the reference's noR_PF template flag) saves one transform compose and one
    H re-expression (about 120 flop, section 8(d): Pin 1.18k -> 1.06k), a leaf body has no child
    inertia / force shift (about 110 flop: 1.06k -> 0.95k)."""
    ny = info.nq + info.nu
    nchild = [0] * info.nb
    for b, p in enumerate(info.parents):
        if b > 0:
            nchild[p] += 1
    ident = ["1", "0", "0", "0", "1", "0", "0", "0", "1"]
    flop = 0.0
    for line in info.text.splitlines():
        tok = line.split()
        if not tok or tok[0] != "body" or tok[3] not in F_EVAL:
            continue
        b = int(tok[1])
        f = F_EVAL[tok[3]]
        if [str(int(float(x))) if float(x) in (0.0, 1.0) else x for x in tok[14:23]] == ident:
            f -= 120.0
        if nchild[b] == 0:
            f -= 110.0
        flop += 5.0 * f
    return flop + 30.0 * ny, 2.0 * 8.0 * ny


def sample_clocks(stop, out):
    """nvidia-smi clocks line of the profiling recipe, sampled during the timed region."""
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                out.append([x.strip() for x in r.stdout.strip().splitlines()[0].split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
            "reasons": reasons}


def run_reference(args, wl, info):
    """The reference's CPU implementation (real Simbody via oracle/_ref/ref_driver)."""
    from _harness import RefDriver, have_ref
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (needs /root/reference at build time)"}))
        return
    ref = RefDriver()
    # bounded sample: per-step cost scales with the body count; aim for a few seconds per bench step
    per_inst_step_s = 4e-5 * max(1, info.nb - 1) / 2.0
    nsteps = 200
    inst = int(max(cores, min(wl["batch"], (3.0 * cores) / (per_inst_step_s * nsteps))))
    inst = (inst // cores) * cores
    q, u = info.random_states(inst, 12345, q_scale=wl["q_scale"])
    y = np.concatenate([q, u], axis=1)
    times = []
    for i in range(args.warmup + args.steps):
        r = ref.bench(info, y, wl["h"], nsteps, cores)
        if i >= args.warmup:
            times.append(r["seconds"])
    total = inst * nsteps * len(times) / sum(times)
    line = {"metric": "instance_steps_per_s", "value": total, "unit": "instance-steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "model": wl["model"], "h": wl["h"], "integrator": "RungeKuttaMerson fixed step"},
            "cpu_baseline": {"value": total, "unit": "instance-steps/s", "cores": cores, "kind": "reference",
                             "sample": "%d instances x %d steps per bench step, one Simbody System per host thread" % (inst, nsteps)},
            "e2e": {"value": total, "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sbk", choices=["sbk", "reference"])
    ap.add_argument("--workload", default="double_pendulum_1M", choices=sorted(WORKLOADS))
    ap.add_argument("--steps-per-launch", type=int, default=0, help="RKM steps per bench step (0 = per-workload default)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-workloads", action="store_true", help="(default behaviour now) also report the other configs in 'workloads'")
    ap.add_argument("--no-extra-workloads", action="store_true", help="report only --workload")
    args = ap.parse_args()

    from _harness import ModelInfo
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch

    if args.impl == "reference":
        import simbody_b200 as sb
        info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
        run_reference(args, wl, info)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keep stdout to the one JSON line
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    import simbody_b200 as sb

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def measure(name, wl, steps, warmup, spl):
        info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
        N = wl["batch"]; ny = info.nq + info.nu
        if not spl:
            # enough RKM steps per launch that one bench step is >= ~20 ms of device work
            spl = {"double_pendulum": 200, "pin_chain": 10, "humanoid30": 10, "branched_tree": 2}[wl["model"]]
        # a non-default stream: handle 0 would mean "library-owned stream" to sbk_batch_create,
        # and torch.cuda.Event only times the stream it is recorded on
        stream = torch.cuda.Stream(device=local)
        topo = sb.Topology(text=info.text)
        bm = sb.BatchedMatter(topo, N, device=local, stream=ctypes.c_void_p(stream.cuda_stream))
        q, u = info.random_states(N, 12345 + rank, q_scale=wl["q_scale"])
        qh = torch.from_numpy(np.ascontiguousarray(q.T)).pin_memory(); uh = torch.from_numpy(np.ascontiguousarray(u.T)).pin_memory()
        qo = torch.empty_like(qh).pin_memory(); uo = torch.empty_like(uh).pin_memory()
        dp = lambda t: ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_double))
        lib = bm.lib
        sb.capi.check(lib, lib.sbk_set_state(bm.handle, dp(qh), dp(uh), None))

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident throughput ("value") ----------------------------------------------
        for _ in range(warmup):
            bm.stepBy(wl["h"], spl)
        barrier()
        launches0 = bm.launchCount()
        stop, samples = threading.Event(), []
        th = threading.Thread(target=sample_clocks, args=(stop, samples)); th.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        kern_ms = []
        e0.record(stream)
        for _ in range(steps):
            bm.stepBy(wl["h"], spl)
        e1.record(stream)
        barrier()
        stop.set(); th.join()
        ms = e0.elapsed_time(e1)
        kern_ms.append(bm.lastKernelMs())
        launches = bm.launchCount() - launches0
        st, nbad = bm.status()

        # ---- end to end through the C ABI with host buffers ----------------------------------------
        for _ in range(2):
            sb.capi.check(lib, lib.sbk_set_state(bm.handle, dp(qh), dp(uh), None)); bm.stepBy(wl["h"], spl)
            sb.capi.check(lib, lib.sbk_get_state(bm.handle, dp(qo), dp(uo), None))
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, steps // 2)
        for _ in range(e2e_steps):
            sb.capi.check(lib, lib.sbk_set_state(bm.handle, dp(qh), dp(uh), None))
            bm.stepBy(wl["h"], spl)
            sb.capi.check(lib, lib.sbk_get_state(bm.handle, dp(qo), dp(uo), None))
        barrier()
        e2e_s = time.perf_counter() - t0

        tmax = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_max, e2e_ms_max = float(tmax[0]), float(tmax[1])
        # end-of-run statistics: the only collectives of the job (NCCL): max error norm, bad-instance count
        from simbody_b200.sharding import reduce_stats
        errn = bm.stepBy(wl["h"], 0, want_err_norm=True)
        max_err, nbad, _ = reduce_stats(float(np.nanmax(errn)), int(nbad), ms, dist if world > 1 else None, device="cuda")
        flop, byts = algorithmic_work(info)
        plan = bm.getPlan()
        # operator form of the same path: System::realize(Acceleration) on the resident state (FULL records)
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        lib.sbk_state_touched(bm.handle); sb.capi.check(lib, lib.sbk_realize_acceleration(bm.handle))
        barrier(); ev0.record(stream)
        nrep = 5
        for _ in range(nrep):
            lib.sbk_state_touched(bm.handle); sb.capi.check(lib, lib.sbk_realize_acceleration(bm.handle))
        ev1.record(stream); barrier()
        realize_per_s = world * N * nrep / (ev0.elapsed_time(ev1) * 1e-3)
        res = {"name": name, "info": info, "N": N, "spl": spl, "ms_per_step": ms_max / steps, "plan": plan,
               "value": world * N * spl * steps / (ms_max * 1e-3),
               "e2e": world * N * spl * e2e_steps / (e2e_ms_max * 1e-3),
               "h2d": 8 * ny * N, "d2h": 8 * ny * N, "launches": launches, "kernel_ms_last": kern_ms[-1],
               "flop_per_inst_step": flop, "bytes_per_inst_step": byts, "realize_per_s": realize_per_s, "clocks": clocks_summary(samples), "nbad": int(nbad), "max_err_norm": max_err}
        bm.close(); topo.close()
        return res

    r = measure(args.workload, wl, args.steps, args.warmup, args.steps_per_launch)

    # FP64 roofline denominator: measured DFMA throughput on this GPU (MEASURED_PEAKS.json has none)
    lib = sb.load_library()
    msd = ctypes.c_double()
    iters = 20000
    sb.capi.check(lib, lib.sbk_dfma_probe(local, 148 * 8, 256, iters, ctypes.byref(msd)))
    fp64_peak_tflops = 2.0 * 8 * iters * 148 * 8 * 256 / (msd.value * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    per_gpu_rate = r["value"] / world
    achieved_tflops = per_gpu_rate * r["flop_per_inst_step"] / 1e12
    roofline = {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved_tflops / fp64_peak_tflops,
                "peak_source": "measured live: sbk_dfma_probe DFMA kernel on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_algorithmic_GBs": per_gpu_rate * r["bytes_per_inst_step"] / 1e9,
                "hbm_peak_GBs": peaks.get("hbm_gbs"),
                "kernel": {1: "tpiKernel<OP_RKM> (thread-per-instance, reversible kinematics, persistent task queue)", 2: "fusedRkmKernel (register-resident)",
                           3: "lpKernel<OP_RKM> (level-parallel)", 4: "glRkmKernel (grid-level-parallel, persistent cooperative grid)"}[r["plan"]],
                "kernel_ms_last_launch": r["kernel_ms_last"],
                "fp64_pipe_active_ncu": NCU_FP64_PIPE_ACTIVE.get(args.workload),
                "hbm_traffic_frac_ncu": (NCU_TRAFFIC_PER_INSTANCE_STEP.get(args.workload, 0) * per_gpu_rate / 1e9 / peaks["hbm_gbs"])
                                        if peaks.get("hbm_gbs") else None,
                # dram__bytes_read+write per launch from the committed ncu captures (profiles/), scaled to this launch
                "traffic": NCU_TRAFFIC_PER_INSTANCE_STEP.get(args.workload, 0) * r["N"] * r["spl"] or None}

    line = {"metric": "instance_steps_per_s", "value": r["value"], "unit": "instance-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "model": wl["model"], "instances_per_gpu": r["N"], "h": wl["h"],
                       "rkm_steps_per_bench_step": r["spl"], "integrator": "RungeKuttaMerson fixed step, 5 evals/step",
                       "l2": "state+cache working set exceeds L2 for every workload but branched_tree; no flush needed"},
            "e2e": {"value": r["e2e"], "unit": "instance-steps/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": roofline,
            "realize_acceleration_per_s": r["realize_per_s"], "non_finite_instances": r["nbad"], "max_err_norm_last_step": r["max_err_norm"]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from _harness import RefDriver, have_ref
        if have_ref():
            cores = os.cpu_count() or 1
            info = r["info"]
            per = 4e-5 * max(1, info.nb - 1) / 2.0
            nst = 200
            inst = int(max(cores, (15.0 * cores) / (per * nst))); inst = (inst // cores) * cores
            q, u = info.random_states(inst, 12345, q_scale=wl["q_scale"])
            rb = RefDriver().bench(info, np.concatenate([q, u], axis=1), wl["h"], nst, cores)
            line["cpu_baseline"] = {"value": rb["instance_steps_per_s"], "unit": "instance-steps/s", "cores": cores,
                                    "kind": "reference", "sample": "%d instances x %d steps, one Simbody System per host thread (%.1f s)" % (inst, nst, rb["seconds"])}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "instance-steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    # the other BASELINE configs ride along in 'workloads' (a few seconds each); only the default workload's line is
    # the contract, so a failure here is reported, not fatal
    if not args.no_extra_workloads and (args.all_workloads or args.workload == "double_pendulum_1M"):
        extra = {}
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            try:
                rr = measure(name, dict(WORKLOADS[name]), max(3, args.steps // 2), args.warmup, 0)
            except Exception as ex:           # noqa: BLE001
                extra[name] = {"error": str(ex)[:200]}
                continue
            tf = rr["value"] / world * rr["flop_per_inst_step"] / 1e12
            extra[name] = {"value": rr["value"], "e2e": rr["e2e"], "ms_per_step": rr["ms_per_step"], "fp64_frac": tf / fp64_peak_tflops,
                           "achieved_tflops": tf, "rkm_steps_per_bench_step": rr["spl"], "instances_per_gpu": rr["N"],
                           "realize_acceleration_per_s": rr["realize_per_s"], "plan": rr["plan"],
                           "fp64_pipe_active_ncu": NCU_FP64_PIPE_ACTIVE.get(name),
                           "hbm_traffic_frac_ncu": (NCU_TRAFFIC_PER_INSTANCE_STEP.get(name, 0) * rr["value"] / world / 1e9 / peaks["hbm_gbs"])
                                                   if peaks.get("hbm_gbs") else None}
        line["workloads"] = extra

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
