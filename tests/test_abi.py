"""CPU tests of the C-ABI library: it loads, exports every symbol include/sbk.h declares, the
host-side topology compiler behaves, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import simbody_b200 as sb
from simbody_b200 import capi
from _harness import ROOT, ModelInfo


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sbk.h")).read()
    declared = set(re.findall(r"\b(sbk_[a-z_A-Z0-9]+)\s*\(", hdr))
    lib = sb.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), "libsbk.so does not export " + name
    assert declared == set(capi.SYMBOLS), (declared ^ set(capi.SYMBOLS))
    assert lib.sbk_version() == 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.BodyDesc) == 8 + 8 + 3 * 8 + 6 * 8 + 12 * 8 + 12 * 8
    assert ctypes.sizeof(capi.ForceDesc) == 16 + 16 + 24 + 24
    assert ctypes.sizeof(capi.RkmOpts) == 24


@pytest.mark.parametrize("name,n,expect", [("double_pendulum", 0, (3, 2, 2, 0, 3)), ("pin_chain", 50, (51, 50, 50, 0, 51)),
                                           ("mixed7", 0, (7, 16, 14, 2, 6)), ("humanoid30", 0, (31, 59, 52, 7, 10)),
                                           ("branched_tree", 1000, (1001, 2000, 1750, 250, 11))])
def test_topology_counts(name, n, expect):
    t = sb.Topology(text=sb.model_text(name, n))
    assert (t.nb, t.nq, t.nu, t.nquat, t.nlevels) == expect
    info = ModelInfo(sb.model_text(name, n))
    assert t.q0 == info.q0 and t.u0 == info.u0
    t.close()


def test_topology_from_descs_and_errors():
    def body(parent, jt, mass=1.0):
        b = capi.BodyDesc(); b.parent = parent; b.joint_type = jt; b.mass = mass
        for i in (0, 4, 8):
            b.X_PF[i] = 1.0; b.X_BM[i] = 1.0
        b.unit_inertia_OB_B[0] = b.unit_inertia_OB_B[1] = b.unit_inertia_OB_B[2] = 1.0
        return b
    g = capi.ForceDesc(); g.kind = 1; g.a = 9.8; g.dir[1] = -1.0
    t = sb.Topology(bodies=[body(-1, 0, 0.0), body(0, 5), body(1, 4), body(1, 1)], forces=[g])
    assert (t.nb, t.nq, t.nu, t.nquat) == (4, 12, 10, 2) and t.level == [0, 1, 2, 2]
    t.close()
    with pytest.raises(sb.SbkError):       # parent after child: not a tree in MobilizedBodyIndex order
        sb.Topology(bodies=[body(-1, 0, 0.0), body(2, 1), body(0, 1)])
    with pytest.raises(sb.SbkError):       # unsupported mobilizer kind
        sb.Topology(bodies=[body(-1, 0, 0.0), body(0, 42)])
    s = capi.ForceDesc(); s.kind = 2; s.body = 1; s.coord = 0; s.a = 1.0
    with pytest.raises(sb.SbkError):       # spring on a quaternion mobilizer (reference quirk, Force.cpp:348)
        sb.Topology(bodies=[body(-1, 0, 0.0), body(0, 4)], forces=[s])


def test_model_text_round_trip():
    for name, n in [("mixed7", 0), ("humanoid30", 0), ("branched_tree", 17)]:
        text = sb.model_text(name, n)
        t = sb.Topology(text=text)
        assert t.nb == ModelInfo(text).nb
        t.close()
    with pytest.raises(sb.SbkError):
        sb.Topology(text="not a model")


def test_no_cpu_fallback():
    """Without a usable GPU, creating a batch must fail loudly (SBK_ERR_CUDA), never compute on the CPU."""
    lib = sb.load_library()
    if lib.sbk_device_count() > 0:
        pytest.skip("a GPU is present")
    t = sb.Topology(text=sb.model_text("double_pendulum"))
    with pytest.raises(sb.SbkError) as e:
        sb.BatchedMatter(t, 4)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)
    t.close()


def _build_facade_smoke():
    import subprocess
    out = os.path.join(ROOT, "build", "facade_smoke")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Iinclude", "-I.", "tests/cpp/facade_smoke.cpp", "-Lsimbody_b200", "-lsbk",
                        "-Wl,-rpath," + os.path.join(ROOT, "simbody_b200"), "-o", out], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def test_cpp_facade_compiles_against_the_c_abi():
    """The header-only C++ facade (BatchedMatter.h: the reference's method names over include/sbk.h) must compile
    and link against libsbk.so without CUDA or Simbody headers."""
    _build_facade_smoke()


@pytest.mark.gpu
def test_cpp_facade_smoke_runs():
    import subprocess
    exe = _build_facade_smoke()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
