"""bench.py's evidence plumbing (CPU): the ncu-derived roofline fields are parsed from the committed captures and printed
only when the capture is of the live kernel at the live per-step duration."""
import os, sys
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


@pytest.mark.parametrize("workload", sorted(bench.WORKLOADS))
def test_committed_capture_parses(workload):
    cap = bench.load_capture(workload)
    assert cap is not None and cap["file"].startswith("profiles/r")
    assert cap["duration_ms"] > 0 and cap["dram_bytes"] > 0 and cap["N"] == bench.WORKLOADS[workload]["batch"]
    assert cap["spl"] and cap["fp64_pipe_pct"] is not None


def test_ncu_fields_are_withheld_on_mismatch():
    w = "humanoid30_64k"; cap = bench.load_capture(w)
    name = cap["kernel"].split("::")[-1].split("(")[0]                    # e.g. tpiKernel<7, 1, 2, 1073741886, 1>
    N, spl = cap["N"], cap["spl"]
    ok = bench.ncu_fields(w, name, cap["duration_ms"], N, spl, 1)
    assert "ncu_withheld" not in ok and ok["traffic"] == pytest.approx(cap["dram_bytes"]) and 0 < ok["fp64_pipe_active_ncu"] < 1
    assert ok["dram_bytes_per_instance_step_ncu"] == pytest.approx(cap["dram_bytes"] / (N * spl))
    slow = bench.ncu_fields(w, name, 1.5 * cap["duration_ms"], N, spl, 1)
    assert "ncu_withheld" in slow and "traffic" not in slow
    other = bench.ncu_fields(w, "tpiKernel<7, 1, 2, 8190, 1>", cap["duration_ms"], N, spl, 1)
    assert "different kernel" in other["ncu_withheld"]
    # twice the steps per launch at the same per-step duration: traffic scales with the steps
    twice = bench.ncu_fields(w, name, 2 * cap["duration_ms"], N, 2 * spl, 1)
    assert twice["traffic"] == pytest.approx(2 * cap["dram_bytes"])


def test_algorithmic_work_matches_survey_counts():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _harness import HostEmu, ModelInfo
    info = ModelInfo(HostEmu().model_text("double_pendulum"))
    flop, byts = bench.algorithmic_work(info)
    assert flop == pytest.approx(10170.0) and byts == 64.0                # C2


def test_reference_arm_contract():
    """`bench.py --impl reference`: the real reference on the host cores, one JSON line with the GPU arm's metric / unit / config
    keys, `impl`, a `cpu_baseline` describing the run and a zero-copy `e2e`; ranks other than 0 print nothing and exit 0."""
    import json, subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _harness import have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "pin_chain50_64k", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "instance_steps_per_s" and d["unit"] == "instance-steps/s" and d["higher_is_better"]
    wl = bench.WORKLOADS["pin_chain50_64k"]
    assert d["config"] == bench.workload_config("pin_chain50_64k", wl, wl["batch"], wl["spl"], "weak")
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
