#!/usr/bin/env python
"""bench.py -- batched RK-Merson instance-steps/s on B200 (BASELINE.json metric).

A "step" of this benchmark is ONE launch of the hot path over the batch: `rkm_steps_per_bench_step`
fixed-size Runge-Kutta-Merson steps (5 derivative evaluations each) for every instance.
  value : whole-job instance-steps/s, state resident in HBM, CUDA events on the launch stream
  e2e   : the same through the public C ABI with HOST (pinned) buffers: H2D of q,u + the launch + D2H of q,u inside
          the timed region, once per bench step (so the copies are amortised over rkm_steps_per_bench_step RKM
          steps -- the state of a resident integrator does not leave the device between steps)
  roofline : algorithmic flop / measured time against the FP64 peak measured live by a DFMA
          probe; the ncu fields (DRAM traffic, FP64 pipe) come from the committed capture profiles/r*_prof_<workload>.txt
          and are printed only when that capture is of the SAME kernel at the same per-step duration (+-10%)
  --impl reference : the reference's own CPU path (oracle/_ref, real Simbody, one System per host thread) on the same
          workload, bounded sample.
One process per GPU under torchrun; instances shard across ranks with no data-path collective (--scaling weak: per-GPU
batch fixed; strong: the workload's batch split over the ranks).  NCCL: timing barrier / max, end-of-run statistics and
one all-gather of the final states.
"""
import argparse
import ctypes
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# workload name -> model, size param, per-GPU batch, h, q scale, RKM steps per launch
WORKLOADS = {
    "double_pendulum_1M": dict(model="double_pendulum", n=0, batch=1048576, h=1e-3, q_scale=3.0, spl=200, t_adapt=0.5),
    # 65536 instances = 512 blocks of 128; 37 steps per launch make 512*37 = 64*296 block-steps: whole rounds of the persistent
    # task queue on 148 SMs x 2 CTAs (no partial last round); the Pin-only kernel runs 3 work groups per SM: 512*111 = 128*444
    "pin_chain50_64k":    dict(model="pin_chain", n=50, batch=65536, h=1e-3, q_scale=1.0, spl=111, t_adapt=0.05),
    "humanoid30_64k":     dict(model="humanoid30", n=0, batch=65536, h=1e-3, q_scale=0.5, spl=37, t_adapt=0.05),
    "branched_tree1000_256": dict(model="branched_tree", n=1000, batch=256, h=5e-4, q_scale=0.5, spl=8),
}
F_EVAL = {"PIN": 1180.0, "SLIDER": 1130.0, "UNIVERSAL": 1710.0, "BALL": 2060.0, "FREE": 3800.0, "WELD": 700.0,
          "TRANSLATION": 1800.0, "CYLINDER": 1650.0, "PLANAR": 1950.0, "GIMBAL": 2300.0}
PLAN_NAMES = {1: "thread-per-instance, body-frame sweeps, fused two-sweep RKM, persistent task queue", 2: "register-resident fused chain",
              3: "level-parallel (CTA per instance)", 4: "grid-level-parallel", 5: "cluster-level-parallel (cluster per 32 instances)"}


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


# ---- committed ncu captures ------------------------------------------------------------------------------------------
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
         "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6, "%": 1.0}


def load_capture(workload):
    """Parse the newest profiles/r<round>_prof_<workload>.txt (written by profiles/summarize_ncu.py on the GPU box from an
    `ncu --set full` report).  -> dict(kernel, duration_ms, dram_bytes, fp64_pipe_pct, N, spl, file) or None."""
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", "r*_prof_%s.txt" % workload)):
        m = re.match(r"r(\d+)([a-z]?)_prof_", os.path.basename(path))
        key = (int(m.group(1)), m.group(2)) if m else (0, "")
        if best is None or key > best[0]:
            best = (key, path)
    if not best:
        return None
    cap = {"file": os.path.relpath(best[1], ROOT), "N": None, "spl": None}
    vals = {}
    for line in open(best[1]):
        if line.startswith("# capture:"):
            for kv in line.split()[2:]:
                k, _, v = kv.partition("=")
                if k == "N":
                    cap["N"] = int(v)
                elif k == "rkm_steps_per_launch":
                    cap["spl"] = int(v)
        elif line.startswith("== kernel:"):
            if "kernel" in cap:
                break                               # first launch in the report only
            cap["kernel"] = line.split(":", 1)[1].strip()
        else:
            t = line.split()
            if len(t) >= 3 and t[0][0].isalpha() and t[1] in _UNIT:
                try:
                    vals[t[0]] = float(t[2]) * _UNIT[t[1]]
                except ValueError:
                    pass
    if "kernel" not in cap or "gpu__time_duration.sum" not in vals:
        return None
    cap["duration_ms"] = vals["gpu__time_duration.sum"]
    cap["dram_bytes"] = vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)
    cap["fp64_pipe_pct"] = vals.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
    return cap


def ncu_fields(workload, live_kernel, live_ms_per_launch, N, spl, plan):
    """ncu-derived roofline fields for a live measurement, or the reason they are withheld."""
    cap = load_capture(workload)
    if cap is None:
        return {"ncu_capture": None}
    out = {"ncu_capture": {"file": cap["file"], "kernel": cap["kernel"], "duration_ms": cap["duration_ms"],
                           "instances": cap["N"], "rkm_steps_per_launch": cap["spl"]}}
    if live_kernel not in cap["kernel"].replace("(int)", ""):
        out["ncu_withheld"] = "capture is of a different kernel than the live run (%s)" % live_kernel
        return out
    capN, capS = cap["N"] or N, cap["spl"] or spl
    per_step_cap = cap["duration_ms"] / (capN * capS)
    per_step_live = live_ms_per_launch / (N * spl)
    out["ncu_capture"]["ms_per_instance_step_vs_live"] = per_step_cap / per_step_live
    if not (0.9 <= per_step_cap / per_step_live <= 1.1) and plan != 2:
        out["ncu_withheld"] = "capture and live per-step durations differ by more than 10%"
        return out
    # the register-resident plan touches DRAM once per launch (read y, write y): its traffic does not scale with the steps
    scale = (N / capN) * (1.0 if plan == 2 else spl / capS)
    out["traffic"] = cap["dram_bytes"] * scale
    out["fp64_pipe_active_ncu"] = None if cap["fp64_pipe_pct"] is None else cap["fp64_pipe_pct"] / 100.0
    out["dram_bytes_per_instance_step_ncu"] = cap["dram_bytes"] / (capN * capS)
    return out


def sample_clocks(stop, out, devices=None):
    """Clocks and throttle reasons of the job's GPUs (the profiling recipe's clocks line), sampled during the timed region.
    In-process through NVML (one light query per GPU every 0.2 s, rank 0 only): eight ranks each spawning nvidia-smi five times a
    second kept the driver busy enough to delay the other ranks' kernel launches by milliseconds."""
    if devices is not None and len(devices) == 0:
        return
    devices = devices or [int(os.environ.get("LOCAL_RANK", "0"))]
    try:
        import pynvml
        pynvml.nvmlInit()
        hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in devices]
        R = pynvml
        bits = [getattr(R, "nvmlClocksEventReasonHwSlowdown", getattr(R, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", getattr(R, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", getattr(R, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                getattr(R, "nvmlClocksEventReasonSwPowerCap", getattr(R, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        reasons_of = getattr(R, "nvmlDeviceGetCurrentClocksEventReasons", None) or R.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            for h in hs:
                try:
                    sm = R.nvmlDeviceGetClockInfo(h, R.NVML_CLOCK_SM); mx = R.nvmlDeviceGetMaxClockInfo(h, R.NVML_CLOCK_SM)
                    m = reasons_of(h)
                    out.append([str(float(sm)), str(float(mx))] + ["Active" if (m & b) else "Not Active" for b in bits])
                except Exception:
                    pass
            stop.wait(0.2)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", ",".join(str(d) for d in devices), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                for line in r.stdout.strip().splitlines():
                    out.append([x.strip() for x in line.split(",")])
        except Exception:
            pass
        stop.wait(0.5)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
            "reasons": reasons}


def workload_config(name, wl, N, spl, scaling):
    """The `config` object: identical keys for the GPU arm and the reference arm."""
    return {"workload": name, "model": wl["model"], "instances_per_gpu": N, "h": wl["h"], "rkm_steps_per_bench_step": spl,
            "integrator": "RungeKuttaMerson fixed step, 5 evals/step", "scaling": scaling,
            "l2": ("no flush needed: register-resident plan, 33 MB of state read once and written once per launch (compute bound)"
                   if name == "double_pendulum_1M" else "no flush needed: state + scratch working set exceeds the 126 MB L2"),
            "e2e_note": "H2D + D2H of the full state once per bench step, i.e. amortised over rkm_steps_per_bench_step RKM steps"}


def cpu_sample_size(info, cores, seconds=12.0, nsteps=200):
    """Instances of the bounded CPU sample: about `seconds` of wall time on `cores` threads at the reference's measured
    single-thread cost (~2e-5 s per body per RKM step), a multiple of the thread count.  Same rule for the reference arm and
    the in-line cpu_baseline."""
    per_inst_step_s = 2e-5 * max(1, info.nb - 1)
    inst = int(max(cores, seconds * cores / (per_inst_step_s * nsteps)))
    return (inst // cores) * cores, nsteps


def run_cpu_reference(info, wl, cores):
    from _harness import RefDriver
    inst, nsteps = cpu_sample_size(info, cores)
    q, u = info.random_states(inst, 12345, q_scale=wl["q_scale"])
    rb = RefDriver().bench(info, np.concatenate([q, u], axis=1), wl["h"], nsteps, cores)
    return rb, inst, nsteps


def run_reference(args, name, wl):
    """The reference's CPU implementation (real Simbody via oracle/_ref/ref_driver): no library of this repository is
    loaded by this arm -- the model text comes from ref_driver itself."""
    from _harness import ModelInfo, RefDriver, have_ref
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (needs /root/reference at build time)"}))
        return
    info = ModelInfo(RefDriver().model_text(wl["model"], wl["n"]))
    times, rate = [], []
    for i in range(args.warmup + args.steps):
        rb, inst, nsteps = run_cpu_reference(info, wl, cores)
        if i >= args.warmup:
            times.append(rb["seconds"]); rate.append(rb["instance_steps_per_s"])
    total = inst * nsteps * len(times) / sum(times)
    N = wl["batch"] if args.scaling == "weak" else wl["batch"] // max(1, args.gpus)
    line = {"metric": "instance_steps_per_s", "value": total, "unit": "instance-steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, wl, N, args.steps_per_launch or wl["spl"], args.scaling),
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
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the batch is per GPU; strong: the batch is the whole job, split over the ranks (BASELINE config 2 as written)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-workloads", action="store_true", help="(default behaviour) also report the other configs in 'workloads'")
    ap.add_argument("--no-extra-workloads", action="store_true", help="report only --workload")
    args = ap.parse_args()

    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch

    if args.impl == "reference":
        run_reference(args, args.workload, wl)
        return

    from _harness import ModelInfo
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keep stdout to the one JSON line
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    import simbody_b200 as sb
    from simbody_b200.sharding import gather_final_states, reduce_stats, shard_range

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = sb.load_library()
    # FP64 roofline denominator: measured DFMA throughput on this GPU (MEASURED_PEAKS.json has none), with the SM clock under it
    stop0, samples0 = threading.Event(), []
    th0 = threading.Thread(target=sample_clocks, args=(stop0, samples0) if rank == 0 else (stop0, samples0, [])); th0.start()
    msd = ctypes.c_double(); iters = 20000; best = 1e30
    for _ in range(8):
        sb.capi.check(lib, lib.sbk_dfma_probe(local, 148 * 8, 256, iters, ctypes.byref(msd))); best = min(best, msd.value)
    stop0.set(); th0.join()
    fp64_peak_tflops = 2.0 * 8 * iters * 148 * 8 * 256 / (best * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def measure(name, wl, steps, warmup, spl, want_gather=False):
        info = ModelInfo(sb.model_text(wl["model"], wl["n"]))
        ny = info.nq + info.nu
        if args.scaling == "strong":
            lo, hi = shard_range(wl["batch"], rank, world); N = hi - lo
        else:
            N = wl["batch"]
        spl = spl or wl["spl"]
        # a non-default stream: handle 0 would mean "library-owned stream" to sbk_batch_create,
        # and torch.cuda.Event only times the stream it is recorded on
        stream = torch.cuda.Stream(device=local)
        topo = sb.Topology(text=info.text)
        bm = sb.BatchedMatter(topo, N, device=local, stream=ctypes.c_void_p(stream.cuda_stream))
        q, u = info.random_states(N, 12345 + rank, q_scale=wl["q_scale"])
        qh = torch.from_numpy(np.ascontiguousarray(q.T)).pin_memory(); uh = torch.from_numpy(np.ascontiguousarray(u.T)).pin_memory()
        qo = torch.empty_like(qh).pin_memory(); uo = torch.empty_like(uh).pin_memory()
        dp = lambda t: ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_double))
        sb.capi.check(lib, lib.sbk_set_state(bm.handle, dp(qh), dp(uh), None))

        def barrier():
            # drain this rank's streams BEFORE the collective: the NCCL barrier kernel otherwise slips in between two queued
            # launches of the persistent integrator kernel (which fills every SM) and holds the next one back until the slowest
            # rank has reached ITS next kernel boundary -- seen at 8 GPUs as one extra kernel time on some ranks
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident throughput ("value") ----------------------------------------------
        for _ in range(warmup):
            bm.stepBy(wl["h"], spl)
        barrier()
        launches0 = bm.launchCount()
        stop, samples = threading.Event(), []
        # rank 0 samples every GPU of the job (one node: local ranks = device indices); the other ranks run no sampler
        th = threading.Thread(target=sample_clocks, args=(stop, samples, list(range(world)) if rank == 0 else [])); th.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            bm.stepBy(wl["h"], spl)
        e1.record(stream)
        barrier()
        stop.set(); th.join()
        ms = e0.elapsed_time(e1)
        kern_ms = bm.lastKernelMs()
        launches = bm.launchCount() - launches0
        st, nbad = bm.status()

        # ---- end to end through the C ABI with pinned host buffers: H2D + launch + D2H per bench step, one sync ----------
        def e2e_step():
            sb.capi.check(lib, lib.sbk_set_state_async(bm.handle, dp(qh), dp(uh), None))
            bm.stepBy(wl["h"], spl)
            sb.capi.check(lib, lib.sbk_get_state_async(bm.handle, dp(qo), dp(uo), None))
            sb.capi.check(lib, lib.sbk_synchronize(bm.handle))        # the step's result is on the host
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, steps // 2)
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0

        tmax = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        per_rank = None
        if world > 1:
            mine = torch.tensor([ms / steps, kern_ms], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = {"ms_per_step": [round(float(x[0]), 3) for x in allr], "kernel_ms_last_launch": [round(float(x[1]), 3) for x in allr]}
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_max, e2e_ms_max = float(tmax[0]), float(tmax[1])
        # end-of-run statistics and (once) the final states: the only data-plane collectives of the job (NCCL)
        errn = bm.stepBy(wl["h"], 0, want_err_norm=True)
        max_err, nbad, _ = reduce_stats(float(np.nanmax(errn)), int(nbad), ms, dist if world > 1 else None, device="cuda")
        ntot = N * world if args.scaling == "weak" else wl["batch"]
        gathered = None
        if want_gather and world > 1:
            if args.scaling == "weak":
                yl = np.concatenate([qo.numpy(), uo.numpy()], axis=0)
                g = gather_final_states(yl, ntot, dist, device="cuda")
            else:
                g = gather_final_states(np.concatenate([qo.numpy(), uo.numpy()], axis=0), ntot, dist, device="cuda")
            gathered = {"shape": list(g.shape), "finite": bool(np.all(np.isfinite(g)))}
        flop, byts = algorithmic_work(info)
        plan = bm.getPlan()
        kname = bm.integratorKernelName()
        # operator form of the same path: System::realize(Acceleration) on the resident state (FULL records)
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        lib.sbk_state_touched(bm.handle); sb.capi.check(lib, lib.sbk_realize_acceleration(bm.handle))
        barrier(); ev0.record(stream)
        nrep = 5
        for _ in range(nrep):
            lib.sbk_state_touched(bm.handle); sb.capi.check(lib, lib.sbk_realize_acceleration(bm.handle))
        ev1.record(stream); barrier()
        realize_per_s = ntot * nrep / (ev0.elapsed_time(ev1) * 1e-3)
        # error-controlled form of the same path (Integrator::stepTo, accuracy 1e-3, every instance with its own step size):
        # accepted and attempted steps per second of this GPU over a short horizon from the same initial states
        adaptive = None
        if wl.get("t_adapt"):
            bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
            st_w, at_w, _ = bm.stepTo(0.1 * wl["t_adapt"], accuracy=1e-3, init_step=wl["h"])      # warm-up
            bm.setState(np.ascontiguousarray(q.T), np.ascontiguousarray(u.T), t=0.0)
            st_a, at_a, _ = bm.stepTo(wl["t_adapt"], accuracy=1e-3, init_step=wl["h"])
            del st_w, at_w                  # sbk_set_state re-initialises the integrator: the counters restart from zero
            ms_a = bm.lastKernelMs()
            adaptive = {"t_final": wl["t_adapt"], "accuracy": 1e-3, "accepted_steps_per_s": float(st_a.sum()) / (ms_a * 1e-3),
                        "attempted_steps_per_s": float(at_a.sum()) / (ms_a * 1e-3), "mean_steps": float(st_a.mean()),
                        "min_steps": int(st_a.min()), "max_steps": int(st_a.max()), "kernel_ms": ms_a, "scope": "this GPU"}
        res = {"name": name, "info": info, "N": N, "ntot": ntot, "spl": spl, "ms_per_step": ms_max / steps, "plan": plan, "kernel": kname,
               "adaptive": adaptive,
               "value": ntot * spl * steps / (ms_max * 1e-3),
               "e2e": ntot * spl * e2e_steps / (e2e_ms_max * 1e-3),
               "h2d": 8 * ny * N, "d2h": 8 * ny * N, "launches": launches, "kernel_ms_last": kern_ms,
               "flop_per_inst_step": flop, "bytes_per_inst_step": byts, "realize_per_s": realize_per_s, "clocks": clocks_summary(samples),
               "nbad": int(nbad), "max_err_norm": max_err, "gathered": gathered, "per_rank": per_rank}
        bm.close(); topo.close()
        return res

    def roofline_of(r):
        per_gpu_rate = r["value"] / world
        achieved = per_gpu_rate * r["flop_per_inst_step"] / 1e12
        roof = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_peak_tflops,
                "peak_source": "measured live: sbk_dfma_probe (8 independent DFMA chains per thread, 148x8 CTAs of 256) on this GPU; "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flop_per_instance_step": r["flop_per_inst_step"], "bytes_per_instance_step": r["bytes_per_inst_step"],
                "hbm_algorithmic_GBs": per_gpu_rate * r["bytes_per_inst_step"] / 1e9, "hbm_peak_GBs": peaks.get("hbm_gbs"),
                "kernel": r["kernel"], "plan": PLAN_NAMES.get(r["plan"], str(r["plan"])), "kernel_ms_last_launch": r["kernel_ms_last"], "traffic": None}
        if r["plan"] == 2:
            roof["frac_note"] = ("algorithmic = the reference's operation count (SURVEY.md 8d); the register-resident kernel skips the "
                                 "multiplications by structural 0 / 1 that count includes and pays less than its 50 flop per sincos, so frac can "
                                 "reach 1.0 while the FP64 pipe (fp64_pipe_active_ncu) is the hardware-side utilisation")
        roof.update(ncu_fields(r["name"], r["kernel"], r["kernel_ms_last"], r["N"], r["spl"], r["plan"]))
        if roof.get("dram_bytes_per_instance_step_ncu") and peaks.get("hbm_gbs"):
            roof["hbm_traffic_frac_ncu"] = roof["dram_bytes_per_instance_step_ncu"] * per_gpu_rate / 1e9 / peaks["hbm_gbs"]
        return roof

    r = measure(args.workload, wl, args.steps, args.warmup, args.steps_per_launch, want_gather=True)
    cfg = workload_config(args.workload, wl, r["N"], r["spl"], args.scaling)
    line = {"metric": "instance_steps_per_s", "value": r["value"], "unit": "instance-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": r["e2e"], "unit": "instance-steps/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": roofline_of(r),
            "dfma_probe": {"tflops": fp64_peak_tflops, "ms": best, "clocks": clocks_summary(samples0)},
            "realize_acceleration_per_s": r["realize_per_s"], "non_finite_instances": r["nbad"], "max_err_norm_last_step": r["max_err_norm"],
            "adaptive": r["adaptive"]}
    if r["gathered"]:
        line["final_states_all_gather"] = r["gathered"]
    if r["per_rank"]:
        line["per_rank"] = r["per_rank"]

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from _harness import have_ref
        if have_ref():
            cores = os.cpu_count() or 1
            rb, inst, nst = run_cpu_reference(r["info"], wl, cores)
            line["cpu_baseline"] = {"value": rb["instance_steps_per_s"], "unit": "instance-steps/s", "cores": cores, "kind": "reference",
                                    "sample": "%d instances x %d steps, one Simbody System per host thread (%.1f s)" % (inst, nst, rb["seconds"])}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "instance-steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    # the other BASELINE configs ride along in 'workloads' (a few seconds each); only the headline workload's line is
    # the contract, so a failure here is reported, not fatal
    if not args.no_extra_workloads and (args.all_workloads or args.workload == "double_pendulum_1M") and args.scaling == "weak":
        extra = {}
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            try:
                rr = measure(name, dict(WORKLOADS[name]), max(3, args.steps // 2), args.warmup, 0)
            except Exception as ex:           # noqa: BLE001
                extra[name] = {"error": str(ex)[:200]}
                continue
            roof = roofline_of(rr)
            extra[name] = {"value": rr["value"], "e2e": rr["e2e"], "ms_per_step": rr["ms_per_step"], "fp64_frac": roof["frac"],
                           "achieved_tflops": roof["achieved"], "rkm_steps_per_bench_step": rr["spl"], "instances_per_gpu": rr["N"],
                           "realize_acceleration_per_s": rr["realize_per_s"], "plan": rr["plan"], "roofline": roof, "adaptive": rr["adaptive"]}
        line["workloads"] = extra

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
