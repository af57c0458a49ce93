"""CPU tests (no GPU): the oracle and the host logic.

  * oracle/sbk_oracle.c (plain-C restatement) against the reference's recorded outputs
    (tests/golden, produced by the UNMODIFIED reference) and against SURVEY.md section 8c values;
  * the kernels' per-body math compiled for the host (tests/hostemu, test-only) against the same;
  * when oracle/_ref is present, live differential runs against the compiled reference and the
    lowering round trip (model -> Simbody system -> lowered model).
"""
import os

import numpy as np
import pytest

from _harness import COracle, HostEmu, ModelInfo, RefDriver, check_sdfast2, have_ref, rel_err, ROOT, random_tree_text as _random_tree_text

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODELS = ["double_pendulum", "pin_chain", "mixed7", "mixed7e", "ugdamp5", "welded8", "cartesian8", "twopoint7", "humanoid30", "branched_tree"]


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not (os.path.exists(os.path.join(ROOT, "build", "libsbk_oracle.so")) and os.path.exists(os.path.join(ROOT, "build", "libsbk_hostemu.so"))):
        g.build()
    return True


@pytest.mark.parametrize("name", MODELS)
def test_c_oracle_matches_reference_golden(built, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = info.split_eval_out(COracle().eval(info, g["eval_in"]))
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-10, (name, k, rel_err(got[k], ref[k]))
    ny = info.nq + info.nu
    ys = COracle().step(info, g["step_in"], float(g["h"]), int(g["nsteps"]))
    assert rel_err(ys[:, :ny], g["step_out"][:, :ny]) < 1e-10


@pytest.mark.parametrize("name", MODELS)
def test_kernel_math_on_host_matches_reference_golden(built, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = info.split_eval_out(HostEmu().eval(info, g["eval_in"]))
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-11, (name, k, rel_err(got[k], ref[k]))
    ny = info.nq + info.nu
    # two-point force elements need every body's ground-frame transform: FULL records (lean=0) is the only integrator form for them
    ys = HostEmu().step(info, g["step_in"], float(g["h"]), int(g["nsteps"]), lean=0 if name == "twopoint7" else 1)
    assert rel_err(ys[:, :ny], g["step_out"][:, :ny]) < 1e-10


@pytest.mark.parametrize("name", MODELS)
def test_c_oracle_extras_match_reference_golden(built, name):
    """calcMobilizerReactionForces and the system Jacobian products of the C restatement against the
    reference's recorded outputs (ref_driver extras)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_extras_out(g["extras_out"])
    got = info.split_extras_out(COracle().extras(info, g["extras_in"]))
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-10, (name, k, rel_err(got[k], ref[k]))


def test_c_oracle_reaction_forces_match_sdfast(built):
    """Known answers that do not come from Simbody: TestMobilizerReactionForces.cpp:408-436."""
    def react(info, q, u):
        co = COracle()
        xin = np.concatenate([q, u, np.zeros((q.shape[0], info.nu + 6 * info.nb))], axis=1)
        ein = np.concatenate([q, u, np.zeros((q.shape[0], 4 * info.nu + 6 * info.nb))], axis=1)
        return info.split_extras_out(co.extras(info, xin))["FM_G"], info.split_eval_out(co.eval(info, ein))["X_GB"]
    check_sdfast2(react)


def test_survey_golden_double_pendulum(built):
    """README double pendulum at q=(0.3,-0.7), u=(1.1,-0.4)."""
    info = ModelInfo(HostEmu().model_text("double_pendulum"))
    inp = np.zeros((1, info.eval_in_stride))
    inp[0, :4] = [0.3, -0.7, 1.1, -0.4]
    inp[0, 4:6] = [0.25, -1.5]      # a for M*a
    inp[0, 6:8] = [0.25, -1.5]      # v for M^-1 v
    inp[0, 8:10] = [0.25, -1.5]     # known udot for the residual
    for impl in (COracle(), HostEmu()):
        o = info.split_eval_out(impl.eval(info, inp))
        assert np.allclose(o["udot"][0], [-2.9050308120918449, 6.3138774504076149], rtol=1e-13)
        assert np.allclose(o["Ma"][0], [-2.5148421872844886, -2.308789453178878], rtol=1e-13)
        assert np.allclose(o["MInvv"][0], [0.85821776207819422, -1.9364183372353365], rtol=1e-13)
        assert np.allclose(o["resid"][0], [-2.978678922095626, -3.0882928547364843], rtol=1e-13)


def test_survey_golden_mixed_fixture(built):
    """mixed Free/Ball/Universal/Pin/Slider tree, xorshift64 state."""
    info = ModelInfo(HostEmu().model_text("mixed7"))
    assert (info.nb, info.nq, info.nu) == (7, 16, 14)
    z = 88172645463325252
    vals = []
    for _ in range(30):
        z ^= (z << 13) & 0xFFFFFFFFFFFFFFFF; z ^= z >> 7; z ^= (z << 17) & 0xFFFFFFFFFFFFFFFF
        vals.append((z >> 11) / 2.0**53 * 2 - 1)
    q, u = np.array(vals[:16]), np.array(vals[16:])
    q[0:4] /= np.linalg.norm(q[0:4]); q[7:11] /= np.linalg.norm(q[7:11])
    assert abs(q[0] - (-0.042694314853977691)) < 1e-15 and abs(u[0] - (-0.99110031744089322)) < 1e-15
    udot_ref = [0.17422288626456839, -0.069420612369921753, -0.26318985600444844, -1.7890383082053554, -10.348226139581147,
                -0.016958149641840503, 0.29386900582164843, -1.5669639987250292, -0.22674220709203174, 2.4868968324497542,
                -0.64761928335628627, -1.2002467096587237, -1.1624387427364304, -0.50966663220061525]
    inp = np.zeros((1, info.eval_in_stride)); inp[0, :16] = q; inp[0, 16:30] = u
    for impl in (COracle(), HostEmu()):
        o = info.split_eval_out(impl.eval(info, inp))
        assert rel_err(o["udot"][0], np.array(udot_ref)) < 1e-12


def test_quaternion_projection_rule(built):
    """Quaternions are normalised only when RMS(|q|-1) > consTol or when forced
    (SimbodyMatterSubsystemRep.cpp:4160); both restatements must agree on when."""
    info = ModelInfo(HostEmu().model_text("mixed7"))
    q, u = info.random_states(4, 11, q_scale=0.3)
    y = np.concatenate([q, u], axis=1)
    ny = info.nq + info.nu
    for kw in (dict(), dict(project_every=1), dict(cons_tol=1e-14)):
        a = HostEmu().step(info, y, 5e-3, 20, **kw); b = COracle().step(info, y, 5e-3, 20, **kw)
        assert rel_err(a[:, :ny], b[:, :ny]) < 1e-10
        assert np.array_equal(a[:, ny + 1], b[:, ny + 1])
    forced = HostEmu().step(info, y, 5e-3, 20, project_every=1)
    assert np.all(forced[:, ny + 1] == 20)
    qn = np.linalg.norm(forced[:, 0:4], axis=1)
    assert np.allclose(qn, 1.0, atol=1e-15)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n", [("double_pendulum", 0), ("pin_chain", 12), ("mixed7", 0), ("mixed7e", 0), ("ugdamp5", 0), ("welded8", 0),
                                    ("cartesian8", 0), ("twopoint7", 0), ("humanoid30", 0), ("branched_tree", 33)])
def test_live_reference_differential(built, name, n):
    emu, ref, co = HostEmu(), RefDriver(), COracle()
    text = emu.model_text(name, n)
    info = ModelInfo(text)
    assert ref.lower(text) == text        # lowering a realized Simbody system reproduces the spec bit for bit
    inp = info.random_eval_input(12, 321, q_scale=0.6)
    r = info.split_eval_out(ref.eval(info, inp))
    for impl, tol in ((emu, 1e-11), (co, 1e-10)):
        g = info.split_eval_out(impl.eval(info, inp))
        for k in r:
            assert rel_err(g[k], r[k]) < tol, (name, k)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_slot_rules_match_reference(built):
    """q/u slots follow MobilizedBodyIndex order with max-nq (RigidBodyNodeSpec.h:81-87)."""
    text = HostEmu().model_text("mixed7")
    info = ModelInfo(text)
    lines = RefDriver().slots(text).splitlines()
    assert lines[0].split() == ["nb", "7", "nq", "16", "nu", "14", "nquat", "2"]
    for b, line in enumerate(lines[1:]):
        t = line.split()
        if b == 0:
            continue
        assert int(t[3]) == info.q0[b] and int(t[7]) == info.u0[b]


def test_plans_agree_on_host(built):
    """FULL (all records) and the register-resident fused plan run the same cores in the same
    order: without FMA contraction (host build) they agree bit for bit.  The integrator (LEAN)
    path recovers each parent transform by inverting the kinematic recurrence in its inward sweep
    (kinReverse), so it agrees with them to a few ulp per body, not bitwise."""
    emu = HostEmu()
    for name, n in [("double_pendulum", 0), ("pin_chain", 9), ("pin_chain", 50), ("mixed7", 0), ("humanoid30", 0), ("branched_tree", 25)]:
        info = ModelInfo(emu.model_text(name, n))
        ny = info.nq + info.nu
        q, u = info.random_states(3, 5, q_scale=0.5)
        y = np.concatenate([q, u], axis=1)
        a = emu.step(info, y, 1e-3, 6, lean=1); b = emu.step(info, y, 1e-3, 6, lean=0)
        assert rel_err(a[:, :ny], b[:, :ny]) < 1e-13, name
        assert np.allclose(a[:, ny], b[:, ny], rtol=1e-3, atol=1e-15) and np.array_equal(a[:, ny + 1], b[:, ny + 1]), name
    info = ModelInfo(emu.model_text("double_pendulum"))
    q, u = info.random_states(8, 6, q_scale=3.0)
    y = np.concatenate([q, u], axis=1)
    assert np.array_equal(emu.step(info, y, 1e-3, 30, lean=0), emu.step(info, y, 1e-3, 30, lean=2))


def test_adaptive_rkm_reproduces_readme_run_on_host(built):
    """BASELINE config C1 (README double pendulum, adaptive RKM, accuracy 1e-3, 20 s): SURVEY.md
    section 8c records 126 steps / 169 attempts and the final state; the kernels' step-size
    controller (host build) must reproduce them exactly."""
    emu = HostEmu()
    info = ModelInfo(emu.model_text("double_pendulum"))
    y = np.array([[0.0, 0.0, 0.0, 5.0]])
    for fused in (False, True):
        o = emu.adaptive(info, y, 20.0, allow_interpolation=False, fused=fused)[0]
        assert (o[4], o[5]) == (126, 169)
        assert np.allclose(o[:4], [-0.69185519245297433, -17.50067206152189, 0.4044972455348419, -4.7180318836946675], rtol=1e-9)
        assert abs(o[7] - 0.137659) < 1e-6 and o[8] == 20.0


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n,tf,qs", [("double_pendulum", 0, 3.0, 2.0), ("mixed7", 0, 1.0, 0.5), ("mixed7e", 0, 1.0, 0.5), ("cartesian8", 0, 1.0, 0.5),
                                           ("humanoid30", 0, 0.3, 0.4)])
def test_adaptive_rkm_matches_live_reference(built, name, n, tf, qs):
    emu, ref = HostEmu(), RefDriver()
    info = ModelInfo(emu.model_text(name, n))
    q, u = info.random_states(5, 3, q_scale=qs)
    y = np.concatenate([q, u], axis=1)
    ny = info.nq + info.nu
    r = ref.adaptive(info, y, tf); e = emu.adaptive(info, y, tf, allow_interpolation=False)
    assert np.array_equal(r[:, ny], e[:, ny]) and np.array_equal(r[:, ny + 1], e[:, ny + 1])     # steps, attempts
    assert rel_err(e[:, :ny], r[:, :ny]) < 1e-9 and np.array_equal(r[:, ny + 4], e[:, ny + 4])


# ---- body-frame ("local") sweeps and the fused two-sweep integrator (sbk_local.cuh, sbk_lrkm.cuh) ------------------------
LOCAL_MODELS = ["double_pendulum", "pin_chain", "mixed7", "humanoid30", "branched_tree"]


@pytest.mark.parametrize("name", LOCAL_MODELS)
def test_body_frame_sweeps_match_reference_golden(built, name):
    """udot / qdot of the body-frame recursion against the reference's recorded realize(Acceleration) outputs, and the two
    integrators built on it (three sweeps per evaluation; fused two sweeps) against the recorded RKM run."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    ny = info.nq + info.nu
    d = HostEmu().deriv(info, g["eval_in"][:, :ny], local=1)
    assert rel_err(d[:, :info.nq], ref["qdot"]) < 1e-13 and rel_err(d[:, info.nq:], ref["udot"]) < 1e-10, name
    for lean in (3, 4):
        ys = HostEmu().step(info, g["step_in"], float(g["h"]), int(g["nsteps"]), lean=lean)
        assert rel_err(ys[:, :ny], g["step_out"][:, :ny]) < 1e-10, (name, lean)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n", [("pin_chain", 50), ("mixed7", 0), ("humanoid30", 0), ("branched_tree", 257)])
def test_body_frame_sweeps_match_live_reference(built, name, n):
    emu, ref = HostEmu(), RefDriver()
    info = ModelInfo(emu.model_text(name, n))
    inp = info.random_eval_input(8, 99, q_scale=0.6)
    r = info.split_eval_out(ref.eval(info, inp))
    d = emu.deriv(info, inp[:, :info.nq + info.nu], local=1)
    assert rel_err(d[:, info.nq:], r["udot"]) < 1e-10 and rel_err(d[:, :info.nq], r["qdot"]) < 1e-13


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_projection_and_norm_options_match_live_reference(built):
    """In-step quaternion projection (AbstractIntegratorRep.cpp:137-208): state AND projection count against the real
    Simbody for the consTol rule, setProjectEveryStep, setUseInfinityNorm, and -- through Integrator::initialize's forced
    projection (Integrator.cpp:367-377) -- a start from unnormalised quaternions.  Both ground-frame and fused body-frame
    integrators, plus the C restatement."""
    emu, ref, co = HostEmu(), RefDriver(), COracle()
    info = ModelInfo(emu.model_text("mixed7")); ny = info.nq + info.nu
    q, u = info.random_states(12, 5, q_scale=0.7)
    y0 = np.concatenate([q, u], axis=1)
    fired = 0
    for kw, h, n in [({}, 2e-2, 20), (dict(cons_tol=1e-6), 2e-2, 20), (dict(cons_tol=1e-12), 2e-2, 20), (dict(project_every=1), 1e-2, 20),
                     (dict(inf_norm=1, cons_tol=1e-12), 2e-2, 20)]:
        r = ref.step(info, y0, h, n, **kw)
        nref = r[:, ny + 2] - 1                                   # minus the forced projection of initialize()
        fired += int(nref.sum())
        for impl, lean in ((emu, 1), (emu, 4), (co, None)):
            o = impl.step(info, y0, h, n, **kw) if lean is None else impl.step(info, y0, h, n, lean=lean, **kw)
            assert rel_err(o[:, :ny], r[:, :ny]) < 1e-10, (kw, lean)
            assert np.array_equal(o[:, ny + 1], nref), (kw, lean, o[:, ny + 1], nref)
    assert fired > 200                                            # the regime really exercises the projection branch
    # a start off the unit sphere: the reference normalises in initialize(); the engine does the same before its first step
    y1 = y0.copy(); y1[:, 0:4] *= 1.3; y1[:, 7:11] *= 0.8
    r = ref.step(info, y1, 1e-2, 10)
    yn = y1.copy()
    for s0 in info.quat_q0:
        yn[:, s0:s0 + 4] /= np.linalg.norm(yn[:, s0:s0 + 4], axis=1, keepdims=True)
    for lean in (1, 4):
        assert rel_err(emu.step(info, yn, 1e-2, 10, lean=lean)[:, :ny], r[:, :ny]) < 1e-10


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,tf,qs", [("double_pendulum", 3.0, 2.0), ("mixed7", 1.0, 0.5), ("humanoid30", 0.3, 0.4)])
def test_fused_body_frame_adaptive_matches_live_reference(built, name, tf, qs):
    emu, ref = HostEmu(), RefDriver()
    info = ModelInfo(emu.model_text(name, 0))
    q, u = info.random_states(5, 3, q_scale=qs)
    y = np.concatenate([q, u], axis=1)
    ny = info.nq + info.nu
    r = ref.adaptive(info, y, tf); e = emu.adaptive(info, y, tf, allow_interpolation=False, fused=2)
    assert np.array_equal(r[:, ny], e[:, ny]) and np.array_equal(r[:, ny + 1], e[:, ny + 1])     # steps, attempts
    assert rel_err(e[:, :ny], r[:, :ny]) < 1e-9 and np.array_equal(r[:, ny + 4], e[:, ny + 4])


def test_non_finite_error_norm_survives_the_infinity_norm(built):
    """A NaN state must give a non-finite error norm in Inf-norm mode too (fmax would drop the NaN and the controller would
    accept the step and grow h): every integrator variant."""
    emu = HostEmu()
    info = ModelInfo(emu.model_text("mixed7")); ny = info.nq + info.nu
    q, u = info.random_states(2, 1, q_scale=0.3)
    y = np.concatenate([q, u], axis=1); y[0, info.nq + 3] = np.nan
    for lean in (0, 1, 3, 4):
        for inf in (0, 1):
            o = emu.step(info, y, 1e-3, 2, inf_norm=inf, lean=lean)
            assert not np.isfinite(o[0, ny]) and np.isfinite(o[1, ny]), (lean, inf, o[:, ny])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_two_point_elements_lower_by_probing(built):
    """Force::TwoPointLinearSpring / TwoPointLinearDamper have no getters: lower_simbody.h identifies bodies, stations and
    constants from Force::calcForceContribution at seeded states; user-entered parameters come back exactly."""
    text = HostEmu().model_text("twopoint7")
    assert RefDriver().lower(text) == text
    assert sum(line.startswith("tp") for line in text.splitlines()) == 3


@pytest.mark.parametrize("model,n", [("branched_tree", 1000), ("branched_tree", 200), ("branched_tree", 37), ("humanoid30", 0),
                                     ("pin_chain", 12), ("mixed7", 0)])
@pytest.mark.parametrize("sched", [(64, 64, 0, 1), (64, 8, 0, 1), (128, 64, 0, 2), (128, 32, 0, 4), (16, 8, 0, 2), (8, 8, 4, 1)])
def test_cluster_task_lists_are_complete_ordered_and_deadlock_free(model, n, sched):
    """Plan 5's schedule is data (topology.cpp: cutTreeForWarps): simulate the warps and their CTA / cluster / cross-cluster
    barriers and check that every body runs once per sweep, after its children (inward) / its parent (outward) with a barrier
    or the same warp in between, and that no warp is left waiting -- for one, two and four clusters per group of instances."""
    emu = HostEmu()
    info = ModelInfo(emu.model_text(model, n))
    nwarps, top, cutw, k = sched
    assert emu.cut_check(info, nwarps, top, cutw, k) == 0


def test_cluster_schedule_shape_of_the_c5_tree():
    """The 1000-body binary tree on 128 warps (4 clusters of 4 CTAs): cut at level 8 (128 subtrees of 7 bodies), the three levels
    above it stay inside the CTAs (__syncthreads), the four top levels (at most 8 bodies wide) run on the first CTA: 14 tasks per
    sweep on the longest list, seven CTA barriers, no cluster barrier, one cross-cluster barrier."""
    emu = HostEmu()
    info = ModelInfo(emu.model_text("branched_tree", 1000))
    assert emu.cut_info(info, 128, 32, 0, 4) == (8, 14, 7, 0, 1)
    assert emu.cut_info(info, 64, 64, 0, 1) == (7, 21, 6, 1, 0)


@pytest.mark.parametrize("model,tf", [("mixed7", 0.3), ("humanoid30", 0.05), ("ugdamp5", 0.3)])
def test_lockstep_adaptive_forms_match_the_plain_ones_on_the_host(model, tf):
    """The error-controlled kernels run a CTA-voting, per-body-lockstep form of the attempt (lRkmAttemptLockstep; the ground-frame
    step with its `mine` / `anyFresh` predicates).  On the host (one thread, the vote is the thread's own flag) both must
    reproduce the plain forms bit for bit: same states, same step and attempt counts."""
    emu = HostEmu()
    info = ModelInfo(emu.model_text(model))
    q, u = info.random_states(6, 11, q_scale=0.6)
    y = np.concatenate([q, u], axis=1)
    for plain, lock in ((2, 3), (0, 4)):
        a = emu.adaptive(info, y, tf, allow_interpolation=False, fused=plain)
        b = emu.adaptive(info, y, tf, allow_interpolation=False, fused=lock)
        assert np.array_equal(a, b), (model, plain, lock, np.abs(a - b).max())
        assert a[:, info.nq + info.nu].min() >= 2                      # several steps were taken


@pytest.mark.parametrize("seed", range(12))
def test_cluster_task_lists_on_random_trees(seed):
    """The schedule builder on irregular trees (leaves above the cut, sibling groups that straddle a CTA's eight warps, several
    subtrees per warp): the simulated run must still process every body once per sweep in dependency order, with no warp left
    waiting, for one, two and four clusters per group."""
    rng = np.random.default_rng(1000 + seed)
    nb = int(rng.integers(20, 700)); max_children = int(rng.integers(1, 5))
    emu = HostEmu()
    info = ModelInfo(_random_tree_text(rng, nb, max_children))
    for sched in [(64, 64, 0, 1), (64, 8, 0, 1), (128, 64, 0, 2), (128, 32, 0, 4), (32, 16, 0, 2), (16, 16, 8, 1), (8, 8, 0, 1)]:
        assert emu.cut_check(info, *sched) == 0, (seed, nb, max_children, sched)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case", ["tree600", "tree150", "humanoid30", "random0", "random1", "random2"])
def test_cluster_plan_on_the_host_matches_the_thread_per_instance_integrator(case):
    """Plan 5 end to end without a GPU: one host thread per warp runs lListStep over the real task lists, with cyclic barriers
    for __syncthreads / barrier.cluster / the cross-cluster barrier.  Every schedule shape must reproduce the plain fused
    integrator (the link flags of subtree walks, CTA-local chains and top levels decide which data ride in a warp's carry)."""
    emu = HostEmu()
    if case.startswith("tree"):
        info = ModelInfo(emu.model_text("branched_tree", int(case[4:])))
    elif case.startswith("random"):
        rng = np.random.default_rng(50 + int(case[6:]))
        info = ModelInfo(_random_tree_text(rng, int(rng.integers(60, 400)), int(rng.integers(2, 4))))
    else:
        info = ModelInfo(emu.model_text(case))
    q, u = info.random_states(2, 21, q_scale=0.4)
    y = np.concatenate([q, u], axis=1)
    ny = info.nq + info.nu
    ref = emu.step(info, y, 5e-4, 2, lean=4)
    for sched in [(64, 64, 0, 1), (128, 32, 0, 4), (128, 64, 0, 2), (32, 16, 0, 2), (16, 16, 0, 1), (8, 8, 0, 1)]:
        got = emu.cluster_step(info, y, 5e-4, 2, *sched)
        assert rel_err(got[:, :ny], ref[:, :ny]) < 1e-13, (case, sched, rel_err(got[:, :ny], ref[:, :ny]))
        assert np.allclose(got[:, ny], ref[:, ny], rtol=1e-9, atol=1e-16), (case, sched)        # error norm: another summation order
        assert np.array_equal(got[:, ny + 1], ref[:, ny + 1])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("seed", range(6))
def test_random_trees_match_live_reference(built, seed):
    """Differential test on irregular random trees of Pin / Universal / Ball bodies (1-4 children per body) against the real
    Simbody: accelerations of the body-frame and ground-frame sweeps and of the C restatement, then a few RKM steps of the fused
    body-frame integrator and of the ground-frame one."""
    rng = np.random.default_rng(7000 + seed)
    emu, ref, co = HostEmu(), RefDriver(), COracle()
    info = ModelInfo(_random_tree_text(rng, int(rng.integers(5, 70)), int(rng.integers(1, 5))))
    inp = info.random_eval_input(4, 100 + seed, q_scale=0.6)
    r = info.split_eval_out(ref.eval(info, inp))
    y = inp[:, :info.nq + info.nu]
    for local in (1, 0):
        d = emu.deriv(info, y, local=local)
        assert rel_err(d[:, info.nq:], r["udot"]) < 1e-10 and rel_err(d[:, :info.nq], r["qdot"]) < 1e-13, (seed, local)
    c = info.split_eval_out(co.eval(info, inp))
    assert rel_err(c["udot"], r["udot"]) < 1e-10 and rel_err(c["X_GB"], r["X_GB"]) < 1e-12
    ny = info.nq + info.nu
    yr = ref.step(info, y, 1e-3, 5)[:, :ny]
    for lean in (4, 1):
        ys = emu.step(info, y, 1e-3, 5, lean=lean)[:, :ny]
        assert rel_err(ys, yr) < 1e-10, (seed, lean, rel_err(ys, yr))
