"""GPU parity tests: the CUDA path (through the C ABI) against the reference.

Two sources of truth:
  * tests/golden/*.npz -- outputs of the UNMODIFIED reference recorded by
    tests/golden/make_golden.py (always available);
  * oracle/_ref/ref_driver -- the compiled reference itself, run live on fresh seeded inputs
    when the prebuilt binary travelled with the snapshot.
Tolerance: north_star asks for 1e-10 relative in double precision; measured agreement is
~1e-15, the tests gate at 1e-11 (accelerations / single step) so that regressions are caught
long before the contract is at risk.
"""
import os

import numpy as np
import pytest

import simbody_b200 as sb
from _harness import ModelInfo, RefDriver, check_sdfast2, have_ref, random_tree_text, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-11
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["double_pendulum", "pin_chain", "mixed7", "mixed7e", "ugdamp5", "welded8", "cartesian8", "twopoint7", "humanoid30", "branched_tree"]


def soa(a):
    return np.ascontiguousarray(np.asarray(a).T)


def run_eval(info, ein, plan=None):
    """Run every operator of the C ABI on the inputs of `ref_driver eval`; returns dict like split_eval_out."""
    n = ein.shape[0]
    nq, nu, nb = info.nq, info.nu, info.nb
    topo = sb.Topology(text=info.text)
    assert (topo.nb, topo.nq, topo.nu, topo.nquat) == (nb, nq, nu, info.nquat)
    bm = sb.BatchedMatter(topo, n)
    if plan is not None:
        bm.setPlan(plan)
    o = nq
    q, u = ein[:, :nq], ein[:, nq:nq + nu]
    o = nq + nu
    a, v, ud, f = (ein[:, o + i * nu:o + (i + 1) * nu] for i in range(4))
    F = ein[:, o + 4 * nu:]
    bm.setState(soa(q), soa(u))
    bm.realizeAcceleration()
    res = {"qdot": bm.getQDot().T, "udot": bm.getUDot().T, "qdotdot": bm.getQDotDot().T, "qerr": bm.getQErr().T,
           "X_GB": bm.getBodyTransforms().reshape(nb * 12, n).T, "V_GB": bm.getBodyVelocities().reshape(nb * 6, n).T,
           "A_GB": bm.getBodyAccelerations().reshape(nb * 6, n).T}
    fm, Fb = bm.getAppliedForces()
    res["fmob_sys"], res["Fbody_sys"] = fm.T, Fb.reshape(nb * 6, n).T
    res["Ma"] = bm.multiplyByM(soa(a)).T
    res["MInvv"] = bm.multiplyByMInv(soa(v)).T
    res["resid"] = bm.calcResidualForceIgnoringConstraints(soa(f), soa(F), soa(ud)).T
    res["resid0"] = bm.calcResidualForceIgnoringConstraints().T
    udot_op, A_op = bm.calcAcceleration(soa(f), soa(F))
    res["udot_op"], res["A_GB_op"] = udot_op.T, A_op.reshape(nb * 6, n).T
    st, nbad = bm.status()
    assert nbad == 0
    bm.close(); topo.close()
    return res


@pytest.mark.parametrize("name", MODELS)
def test_eval_matches_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = run_eval(info, g["eval_in"])
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))


@pytest.mark.parametrize("name", MODELS)
def test_rkm_step_matches_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    y0, yref = g["step_in"], g["step_out"]
    n, ny = y0.shape[0], info.nq + info.nu
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    err = bm.stepBy(float(g["h"]), int(g["nsteps"]), want_err_norm=True)
    q, u, t = bm.getState()
    assert np.all(np.isfinite(err))
    assert np.allclose(t, float(g["h"]) * int(g["nsteps"]), rtol=1e-12)
    got = np.concatenate([q.T, u.T], axis=1)
    assert rel_err(got, yref[:, :ny]) < 1e-10, (name, rel_err(got, yref[:, :ny]))
    s = bm.stats()
    assert s["steps_taken"] == n * int(g["nsteps"]) and s["realizations"] == 5 * n * int(g["nsteps"])
    bm.close(); topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n,count", [("double_pendulum", 0, 512), ("pin_chain", 50, 96), ("mixed7", 0, 256),
                                          ("humanoid30", 0, 128), ("branched_tree", 60, 64)])
def test_eval_matches_live_reference(name, n, count):
    info = ModelInfo(sb.model_text(name, n))
    ein = info.random_eval_input(count, 4242, q_scale=2.0 if name == "double_pendulum" else 0.7)
    ref = info.split_eval_out(RefDriver().eval(info, ein))
    got = run_eval(info, ein)
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n,count,h,nsteps", [("double_pendulum", 0, 256, 1e-3, 100), ("pin_chain", 50, 32, 1e-3, 10),
                                                   ("mixed7", 0, 64, 2e-3, 50), ("humanoid30", 0, 64, 1e-3, 20)])
def test_rkm_matches_live_reference(name, n, count, h, nsteps):
    info = ModelInfo(sb.model_text(name, n))
    q, u = info.random_states(count, 99, q_scale=0.5)
    y0 = np.concatenate([q, u], axis=1)
    yref = RefDriver().step(info, y0, h, nsteps)[:, :info.nq + info.nu]
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, count)
    bm.setStateAoS(q, u)
    bm.stepBy(h, nsteps)
    qg, ug = bm.getStateAoS()
    assert rel_err(np.concatenate([qg, ug], axis=1), yref) < 1e-10
    bm.close(); topo.close()


def test_mass_matrix_invariants_large_batch():
    """Size-independent properties at a large batch (reference TestMassMatrix.cpp:670-787):
    M^-1 (M v) = v, and inverse dynamics o forward dynamics = identity."""
    info = ModelInfo(sb.model_text("humanoid30"))
    n = 4096
    q, u = info.random_states(n, 7, q_scale=0.5)
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
    bm.setState(soa(q), soa(u))
    bm.realizeVelocityKinematics()
    rng = np.random.default_rng(3)
    v = rng.uniform(-1, 1, size=(info.nu, n))
    back = bm.multiplyByMInv(bm.multiplyByM(v))
    assert rel_err(back, v) < 1e-9
    f = rng.uniform(-1, 1, size=(info.nu, n)); F = rng.uniform(-1, 1, size=(info.nb * 6, n))
    udot, _ = bm.calcAcceleration(f, F)
    resid = bm.calcResidualForceIgnoringConstraints(f, F, udot)
    assert float(np.max(np.abs(resid))) < 1e-8
    bm.close(); topo.close()


def test_stage_and_argument_errors():
    info = ModelInfo(sb.model_text("double_pendulum"))
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, 8)
    with pytest.raises(sb.SbkError) as e:      # operator before realize: SimTK_STAGECHECK analogue
        bm.multiplyByM(np.zeros((2, 8)))
    assert e.value.code == 2
    with pytest.raises(sb.SbkError) as e:      # wrong length: SimTK_APIARGCHECK analogue
        bm.realizePositionKinematics(); bm.multiplyByM(np.zeros((3, 8)))
    assert e.value.code == 1
    bm.close(); topo.close()


def test_fused_plan_matches_generic_plan():
    """The register-resident fused plan calls the same cores in the same order as the
    thread-per-instance plan (bit-identical in the host build, tests/test_oracle.py); on the GPU
    nvcc contracts FMAs differently in the two inlining contexts, so agreement is to rounding."""
    g = np.load(os.path.join(GOLDEN, "double_pendulum.npz"))
    info = ModelInfo(str(g["text"]))
    n = 4096
    q, u = info.random_states(n, 17, q_scale=3.0)
    out = {}
    for plan in (1, 2):
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
        bm.setPlan(plan); assert bm.getPlan() == plan
        bm.setState(soa(q), soa(u), t=0.0)
        err = bm.stepBy(1e-3, 40, want_err_norm=True)
        qq, uu, t = bm.getState()
        out[plan] = (qq, uu, err, t)
        bm.close(); topo.close()
    for a, b in zip(out[1][:2], out[2][:2]):
        assert rel_err(a, b) < 1e-11
    assert np.allclose(out[1][3], out[2][3], rtol=1e-12)
    # error norms are differences of nearly equal numbers: only their magnitude is comparable
    assert np.all(np.isfinite(out[2][2])) and np.allclose(out[1][2], out[2][2], rtol=0.5, atol=1e-13)
    topo = sb.Topology(text=sb.model_text("humanoid30")); bm = sb.BatchedMatter(topo, 8)
    assert bm.getPlan() == 1
    with pytest.raises(sb.SbkError):
        bm.setPlan(2)
    bm.close(); topo.close()


@pytest.mark.parametrize("name", ["mixed7", "mixed7e", "welded8", "cartesian8", "twopoint7", "humanoid30", "branched_tree"])
def test_level_parallel_plan_matches_golden(name):
    """Plan 3 (CTA per instance, threads over the bodies of a level) against the reference."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = run_eval(info, g["eval_in"], plan=3)
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))
    y0, yref = g["step_in"], g["step_out"]
    n, ny = y0.shape[0], info.nq + info.nu
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n); bm.setPlan(3)
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    err = bm.stepBy(float(g["h"]), int(g["nsteps"]), want_err_norm=True)
    q, u, t = bm.getState()
    assert np.all(np.isfinite(err))
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref[:, :ny]) < 1e-10
    bm.close(); topo.close()


@pytest.mark.parametrize("name", ["mixed7", "mixed7e", "welded8", "cartesian8", "twopoint7", "humanoid30", "branched_tree"])
def test_grid_level_parallel_plan_matches_golden(name):
    """Plan 4 (persistent grid, work items = body of a level x warp of instances, grid barriers between
    levels) against the reference: fixed-step RKM, plus agreement with plan 1 on a batch that is not a
    multiple of the warp size and spans several record blocks."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = run_eval(info, g["eval_in"], plan=4)            # the API operations run as grid-level sweeps too
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))
    y0, yref = g["step_in"], g["step_out"]
    n, ny = y0.shape[0], info.nq + info.nu
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n); bm.setPlan(4); assert bm.getPlan() == 4
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    err = bm.stepBy(float(g["h"]), int(g["nsteps"]), want_err_norm=True)
    q, u, t = bm.getState()
    assert np.all(np.isfinite(err)) and np.allclose(t, float(g["h"]) * int(g["nsteps"]))
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref[:, :ny]) < 1e-10
    bm.close()
    nb = 300
    qq, uu = info.random_states(nb, 31, q_scale=0.5)
    out = {}
    for plan in (1, 4):
        bm = sb.BatchedMatter(topo, nb); bm.setPlan(plan)
        bm.setState(soa(qq), soa(uu), t=0.0)
        e = bm.stepBy(1e-3, 3, want_err_norm=True)
        a, b, _ = bm.getState()
        out[plan] = (a, b, e)
        bm.close()
    assert rel_err(out[4][0], out[1][0]) < 1e-10 and rel_err(out[4][1], out[1][1]) < 1e-10, (rel_err(out[4][0], out[1][0]), rel_err(out[4][1], out[1][1]))
    assert np.allclose(out[4][2], out[1][2], rtol=1e-3, atol=1e-14)
    topo.close()


def test_auto_plan_selection():
    for name, n, batch, plan in [("double_pendulum", 0, 1024, 2), ("pin_chain", 50, 1024, 1), ("humanoid30", 0, 1024, 1),
                                 ("branched_tree", 1000, 64, 5), ("branched_tree", 1000, 65536 // 64 * 16, 1)]:
        topo = sb.Topology(text=sb.model_text(name, n))
        if name == "branched_tree" and batch > 64:
            batch = 16384       # large batches of wide trees go thread-per-instance
        bm = sb.BatchedMatter(topo, batch)
        assert bm.getPlan() == plan, (name, batch, bm.getPlan())
        bm.close(); topo.close()


def test_adaptive_rkm_readme_double_pendulum():
    """BASELINE config C1 on the GPU: README double pendulum, adaptive RKM (accuracy 1e-3) to 20 s.
    The reference takes 126 steps / 169 attempts.  The step-size sequence
    is chaotic in the last digits of the error norm, so the GPU run must reproduce the counts within a
    few steps and every instance must land exactly on t = 20; both plans are exercised."""
    info = ModelInfo(sb.model_text("double_pendulum"))
    n = 256
    q = np.zeros((2, n)); u = np.zeros((2, n)); u[1, :] = 5.0
    for plan in (2, 1):
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n); bm.setPlan(plan)
        bm.setState(q, u, t=0.0)
        steps, att, last = bm.stepTo(20.0)
        qq, uu, t = bm.getState()
        assert np.all(t == 20.0)
        assert np.all(steps == steps[0]) and np.all(att == att[0])          # identical instances, identical history
        assert abs(int(steps[0]) - 126) <= 3 and abs(int(att[0]) - 169) <= 4, (plan, steps[0], att[0])
        st, nbad = bm.status(); assert nbad == 0
        bm.close(); topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_adaptive_rkm_short_horizon_matches_live_reference():
    """Short horizons (before chaos amplifies rounding): state, step and attempt counts against the reference."""
    for name, tf, qs, plan in [("double_pendulum", 1.0, 2.0, 2), ("double_pendulum", 1.0, 2.0, 1), ("humanoid30", 0.2, 0.4, 1)]:
        info = ModelInfo(sb.model_text(name))
        nI = 64
        qv, uv = info.random_states(nI, 21, q_scale=qs)
        ref = RefDriver().adaptive(info, np.concatenate([qv, uv], axis=1), tf)
        ny = info.nq + info.nu
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI); bm.setPlan(plan)
        bm.setState(soa(qv), soa(uv), t=0.0)
        steps, att, last = bm.stepTo(tf)
        qq, uu, t = bm.getState()
        same = (steps == ref[:, ny]) & (att == ref[:, ny + 1])
        assert same.mean() > 0.9, (name, plan, same.mean())
        got = np.concatenate([qq.T, uu.T], axis=1)
        assert rel_err(got[same], ref[same][:, :ny]) < 1e-8, (name, plan)
        assert np.all(t == tf)
        bm.close(); topo.close()


@pytest.mark.parametrize("name", MODELS)
def test_energy_matches_golden(name):
    """Kinetic / potential energy against MultibodySystem::calcKineticEnergy / calcPotentialEnergy."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ein = g["eval_in"]; n = ein.shape[0]
    for plan in ([None, 3] if name in ("mixed7", "branched_tree") else [None]):
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
        if plan:
            bm.setPlan(plan)
        bm.setState(soa(ein[:, :info.nq]), soa(ein[:, info.nq:info.nq + info.nu]))
        bm.realizeVelocityKinematics()
        ke, pe = bm.calcEnergy()
        assert rel_err(ke, g["energy"][:, 0]) < 1e-12 and rel_err(pe, g["energy"][:, 1]) < 1e-12
        bm.close(); topo.close()


def test_energy_conservation_over_a_run():
    """Undamped systems conserve KE+PE (reference Simbody/tests/TestForces.cpp:244-288)."""
    for name, n, h, steps in [("double_pendulum", 0, 5e-4, 4000), ("pin_chain", 10, 5e-4, 1000)]:
        info = ModelInfo(sb.model_text(name, n))
        nI = 128
        q, u = info.random_states(nI, 5, q_scale=1.0)
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI)
        bm.setState(soa(q), soa(u))
        bm.realizeVelocityKinematics(); ke0, pe0 = bm.calcEnergy()
        bm.stepBy(h, steps)
        bm.realizeVelocityKinematics(); ke1, pe1 = bm.calcEnergy()
        drift = np.abs((ke1 + pe1) - (ke0 + pe0)) / np.maximum(1.0, np.abs(ke0 + pe0))
        assert drift.max() < 1e-7, (name, drift.max())
        bm.close(); topo.close()


def run_extras(info, xin, plan=None):
    n = xin.shape[0]; nq, nu, nb = info.nq, info.nu, info.nb
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
    if plan is not None:
        bm.setPlan(plan)
    bm.setState(soa(xin[:, :nq]), soa(xin[:, nq:nq + nu]))
    with pytest.raises(sb.SbkError):
        bm.calcMobilizerReactionForces()            # Stage::Acceleration not realized yet
    bm.realizeAcceleration()
    res = {"FM_G": bm.calcMobilizerReactionForces().reshape(nb * 6, n).T,
           "Jv": bm.multiplyBySystemJacobian(soa(xin[:, nq + nu:nq + 2 * nu])).reshape(nb * 6, n).T,
           "JtF": bm.multiplyBySystemJacobianTranspose(soa(xin[:, nq + 2 * nu:])).T,
           "CBI": bm.calcCompositeBodyInertias().reshape(nb * 10, n).T[:, 10:],
           "X_GB": bm.getBodyTransforms().reshape(nb * 12, n).T}
    bm.close(); topo.close()
    return res


@pytest.mark.parametrize("name", MODELS)
def test_reactions_and_jacobian_match_golden(name):
    """calcMobilizerReactionForces, multiplyBySystemJacobian[Transpose] against the reference's recorded outputs."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_extras_out(g["extras_out"])
    got = run_extras(info, g["extras_in"])
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))
    if name in ("mixed7", "mixed7e", "welded8", "cartesian8", "humanoid30", "branched_tree"):
        for plan in (3, 4):                                    # level-parallel record layout; grid-level realize
            gotp = run_extras(info, g["extras_in"], plan=plan)
            for k in ref:
                assert rel_err(gotp[k], ref[k]) < TOL, (name, k, "plan %d" % plan)


def test_reaction_forces_match_sdfast_known_answers():
    """TestMobilizerReactionForces.cpp:408-436: values generated by SD/FAST, independent of Simbody."""
    def react(info, q, u):
        xin = np.concatenate([q, u, np.zeros((q.shape[0], info.nu + 6 * info.nb))], axis=1)
        r = run_extras(info, xin)
        return r["FM_G"], r["X_GB"]
    check_sdfast2(react)


def test_jacobian_transpose_is_adjoint_of_jacobian():
    """Size-independent property at a full-size batch: <J v, F> == <v, ~J F> for every instance."""
    info = ModelInfo(sb.model_text("humanoid30"))
    n = 4096
    x = info.random_extras_input(n, 77, q_scale=0.5)
    r = run_extras(info, x)
    v = x[:, info.nq + info.nu:info.nq + 2 * info.nu]; F = x[:, info.nq + 2 * info.nu:]
    lhs = np.sum(r["Jv"] * F, axis=1); rhs = np.sum(v * r["JtF"], axis=1)
    assert np.max(np.abs(lhs - rhs) / (1 + np.abs(lhs))) < 1e-12


@pytest.mark.parametrize("n", [1, 33, 129, 257])
def test_ragged_batch_sizes(n):
    """Batches that are not multiples of the warp / record-block size (1, 33, 129, 257 instances): every instance
    must reproduce the reference's fixture values (the fixture states are tiled over the batch)."""
    g = np.load(os.path.join(GOLDEN, "humanoid30.npz"))
    info = ModelInfo(str(g["text"]))
    ny = info.nq + info.nu
    reps = -(-n // g["eval_in"].shape[0])
    ein = np.tile(g["eval_in"], (reps, 1))[:n]; eref = np.tile(g["eval_out"], (reps, 1))[:n]
    got = run_eval(info, ein); ref = info.split_eval_out(eref)
    for k in ("udot", "A_GB", "MInvv", "resid"):
        assert rel_err(got[k], ref[k]) < TOL, (n, k)
    y0 = np.tile(g["step_in"], (reps, 1))[:n]; yref = np.tile(g["step_out"], (reps, 1))[:n]
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n)
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    bm.stepBy(float(g["h"]), int(g["nsteps"]))
    q, u, t = bm.getState()
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref[:, :ny]) < 1e-10
    st, nbad = bm.status(); assert nbad == 0
    bm.close(); topo.close()


# ============================================================================================================
# Round-2 parity additions: the benched configurations at their real sizes, the projection / norm options against the
# live reference, Integrator::initialize's projection, the per-instance status word.
# ============================================================================================================
def _with_env(**kw):
    """Context manager: environment overrides read at sbk_batch_create (e.g. SBK_NOLOCAL=1 -> ground-frame integrator)."""
    import contextlib

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kw}
        os.environ.update({k: str(v) for k, v in kw.items()})
        try:
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return cm()


@pytest.mark.parametrize("plan", [1, 3, 4, 5])
def test_branched_tree_1000_matches_golden(plan):
    """BASELINE config 5 at its benched size (1000 bodies, 11 levels, width 489; tables too large to stage): every operator
    and a fixed-step RKM run against the reference's recorded outputs, through every plan that can run it."""
    g = np.load(os.path.join(GOLDEN, "branched_tree1000.npz"))
    info = ModelInfo(str(g["text"]))
    assert info.nb == 1001
    ref = info.split_eval_out(g["eval_out"])
    got = run_eval(info, g["eval_in"], plan=plan)
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-10, (plan, k, rel_err(got[k], ref[k]))
    y0, yref = g["step_in"], g["step_out"]
    n, ny = y0.shape[0], info.nq + info.nu
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n); bm.setPlan(plan); assert bm.getPlan() == plan
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    bm.stepBy(float(g["h"]), int(g["nsteps"]))
    q, u, t = bm.getState()
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref[:, :ny]) < 1e-10, plan
    assert bm.stats()["q_projections"] == int(yref[:, ny + 2].sum())      # only the forced projection of initialize()
    bm.close(); topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,n,batch,h,nsteps,qs", [("double_pendulum", 0, 1048576, 1e-3, 200, 3.0), ("pin_chain", 50, 65536, 1e-3, 111, 1.0),
                                                      ("humanoid30", 0, 65536, 1e-3, 37, 0.5), ("branched_tree", 1000, 256, 5e-4, 8, 0.5)])
def test_full_size_batches_match_live_reference(name, n, batch, h, nsteps, qs):
    """The four BASELINE batches at the size and the steps-per-launch bench.py runs (several rounds of the persistent
    task queue / the whole cooperative grid): a strided sample of 64+ instances -- first, last, spread over every
    block round -- against the real Simbody on the same states."""
    info = ModelInfo(sb.model_text(name, n))
    ny = info.nq + info.nu
    q, u = info.random_states(batch, 12345, q_scale=qs)
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, batch)
    bm.setState(soa(q), soa(u), t=0.0)
    bm.stepBy(h, nsteps)
    qg, ug, t = bm.getState()
    st, nbad = bm.status(); assert nbad == 0
    idx = np.unique(np.concatenate([np.linspace(0, batch - 1, 64).astype(int), [0, 1, 127, 128, batch - 129, batch - 2, batch - 1]]))
    yref = RefDriver().step(info, np.concatenate([q[idx], u[idx]], axis=1), h, nsteps)[:, :ny]
    got = np.concatenate([qg.T[idx], ug.T[idx]], axis=1)
    assert rel_err(got, yref) < 1e-10, (name, rel_err(got, yref))
    assert np.allclose(t, h * nsteps, rtol=1e-12)
    bm.close(); topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("variant", ["local", "ground", "plan3", "plan4", "plan5"])
def test_projection_options_match_live_reference(variant):
    """In-step quaternion projection in a regime where it FIRES (AbstractIntegratorRep.cpp:137-208): state and the number of
    projections (initialize()'s forced one included, Integrator.cpp:367-377) against the real Simbody, for the consTol rule,
    setProjectEveryStep and setUseInfinityNorm; then a start from unnormalised quaternions."""
    info = ModelInfo(sb.model_text("mixed7")); ny = info.nq + info.nu
    nI = 48
    q, u = info.random_states(nI, 5, q_scale=0.7)
    y0 = np.concatenate([q, u], axis=1)
    env = dict(SBK_NOLOCAL=1) if variant == "ground" else {}
    plan = {"plan3": 3, "plan4": 4, "plan5": 5}.get(variant, 1)
    fired = 0
    cases = [(dict(), {}, 2e-2, 20), (dict(cons_tol=1e-12), dict(constraint_tol=1e-12), 2e-2, 20),
             (dict(project_every=1), dict(project_every_step=True), 1e-2, 20),
             (dict(inf_norm=1, cons_tol=1e-12), dict(use_infinity_norm=True, constraint_tol=1e-12), 2e-2, 20)]
    with _with_env(**env):
        for rkw, gkw, h, n in cases:
            r = RefDriver().step(info, y0, h, n, **rkw)
            topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI); bm.setPlan(plan)
            bm.setState(soa(q), soa(u), t=0.0)
            bm.stepBy(h, n, **gkw)
            qg, ug, _ = bm.getState()
            assert rel_err(np.concatenate([qg.T, ug.T], axis=1), r[:, :ny]) < 1e-10, (variant, rkw)
            assert bm.stats()["q_projections"] == int(r[:, ny + 2].sum()), (variant, rkw, bm.stats(), r[:, ny + 2].sum())
            fired += int(r[:, ny + 2].sum()) - nI
            bm.close(); topo.close()
        assert fired > 500
        y1 = y0.copy(); y1[:, 0:4] *= 1.3; y1[:, 7:11] *= 0.8           # off the unit sphere: initialize() projects first
        r = RefDriver().step(info, y1, 1e-2, 10)
        topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI); bm.setPlan(plan)
        bm.setState(soa(y1[:, :info.nq]), soa(y1[:, info.nq:]), t=0.0)
        bm.stepBy(1e-2, 10)
        qg, ug, _ = bm.getState()
        assert rel_err(np.concatenate([qg.T, ug.T], axis=1), r[:, :ny]) < 1e-10, variant
        bm.close(); topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("name,tf,qs,kw", [("mixed7", 0.5, 0.5, dict(inf_norm=1)), ("humanoid30", 0.2, 0.4, dict(inf_norm=1)),
                                           ("mixed7", 0.5, 0.5, dict(cons_tol=1e-9)), ("humanoid30", 0.2, 0.4, dict())])
def test_adaptive_options_match_live_reference(name, tf, qs, kw):
    """Error-controlled stepping with setUseInfinityNorm / a tight constraint tolerance: step and attempt counts, projection
    counts and the state against the real Simbody (fused body-frame integrator and the ground-frame one)."""
    info = ModelInfo(sb.model_text(name)); ny = info.nq + info.nu
    nI = 32
    qv, uv = info.random_states(nI, 21, q_scale=qs)
    ref = RefDriver().adaptive(info, np.concatenate([qv, uv], axis=1), tf, allow_interpolation=False, **kw)
    gkw = dict(use_infinity_norm=bool(kw.get("inf_norm", 0)))
    if "cons_tol" in kw:
        gkw["constraint_tol"] = kw["cons_tol"]
    for env in ({}, dict(SBK_NOLOCAL=1)):
        with _with_env(**env):
            topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI)
            bm.setState(soa(qv), soa(uv), t=0.0)
            steps, att, last = bm.stepTo(tf, **gkw)
            qq, uu, t = bm.getState()
            same = (steps == ref[:, ny]) & (att == ref[:, ny + 1])
            assert same.mean() > 0.9, (name, env, same.mean())
            got = np.concatenate([qq.T, uu.T], axis=1)
            assert rel_err(got[same], ref[same][:, :ny]) < 1e-8, (name, env)
            if same.all():
                assert bm.stats()["q_projections"] == int(ref[:, ny + 5].sum()), (name, env)
                s = bm.stats()
                assert s["steps_taken"] == int(ref[:, ny].sum()) and s["realizations"] == int(ref[:, ny].sum() + 4 * ref[:, ny + 1].sum())
            assert np.all(t == tf)
            bm.close(); topo.close()


def test_adaptive_runs_in_the_auto_plan_of_wide_trees():
    """stepTo() must work out of the box on the models the auto plan sends to plan 4 (per-instance step histories run the
    thread-per-instance kernel on the shared record layout)."""
    info = ModelInfo(sb.model_text("branched_tree", 200))
    nI = 8
    q, u = info.random_states(nI, 3, q_scale=0.4)
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI); assert bm.getPlan() == 5
    bm.setState(soa(q), soa(u), t=0.0)
    steps, att, last = bm.stepTo(0.02)
    qa, ua, t = bm.getState()
    assert np.all(t == 0.02) and np.all(steps >= 1)
    bm.close()
    bm = sb.BatchedMatter(topo, nI); bm.setPlan(1)
    bm.setState(soa(q), soa(u), t=0.0)
    s1, a1, _ = bm.stepTo(0.02)
    q1, u1, _ = bm.getState()
    assert np.array_equal(s1, steps) and np.array_equal(a1, att) and rel_err(qa, q1) < 1e-12
    bm.close()
    # the CTA-per-instance plan borrows the thread-per-instance layout for the call and keeps working afterwards
    bm = sb.BatchedMatter(topo, nI); bm.setPlan(3)
    bm.setState(soa(q), soa(u), t=0.0)
    s3, a3, _ = bm.stepTo(0.02)
    q3, u3, _ = bm.getState()
    assert bm.getPlan() == 3 and np.array_equal(s3, steps) and np.array_equal(a3, att) and rel_err(q3, q1) < 1e-12
    bm.stepBy(5e-4, 2); bm.realizeAcceleration()
    assert np.all(np.isfinite(bm.getUDot()))
    bm.close(); topo.close()


@pytest.mark.parametrize("variant", ["local", "ground", "plan3", "plan4", "plan5", "fused2"])
def test_status_word_flags_bad_instances_only(variant):
    """Per-instance status word: bit 0 = non-finite error norm (a NaN state, RMS and Inf norm), bit 1 = singular joint-space
    inertia D; healthy instances of the same batch stay 0 and finite; a new state clears the word."""
    if variant == "fused2":
        info = ModelInfo(sb.model_text("double_pendulum")); plan = 2
    else:
        info = ModelInfo(sb.model_text("mixed7")); plan = {"plan3": 3, "plan4": 4, "plan5": 5}.get(variant, 1)
    nI = 200
    q, u = info.random_states(nI, 9, q_scale=0.3)
    bad = [3, 131, 199]
    with _with_env(**(dict(SBK_NOLOCAL=1) if variant == "ground" else {})):
        for inf in (False, True):
            ub = u.copy(); ub[bad, 1] = np.nan
            topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, nI); bm.setPlan(plan)
            bm.setState(soa(q), soa(ub), t=0.0)
            err = bm.stepBy(1e-3, 3, use_infinity_norm=inf, want_err_norm=True)
            st, nbad = bm.status()
            assert nbad == len(bad) and sorted(np.nonzero(st)[0].tolist()) == bad and np.all(st[bad] & 1), (variant, inf, np.nonzero(st)[0])
            good = np.setdiff1d(np.arange(nI), bad)
            assert np.all(np.isfinite(err[good])) and not np.any(np.isfinite(err[bad]))
            qg, ug, _ = bm.getState()
            assert np.all(np.isfinite(qg[:, good])) and np.all(np.isfinite(ug[:, good]))
            bm.setState(soa(q), soa(u), t=0.0)                     # a new state starts a new history
            bm.stepBy(1e-3, 1)
            st, nbad = bm.status(); assert nbad == 0
            bm.close(); topo.close()
    if variant in ("local", "ground", "plan3", "plan4", "plan5"):
        # a leaf Pin body with no inertia about its own axis: D = ~H P H = 0 (RigidBodyNodeSpec.cpp:293 inverts it)
        text = sb.model_text("double_pendulum").splitlines()
        tok = text[5].split(); assert tok[0] == "body" and tok[1] == "2"
        tok[8:14] = ["0"] * 6; tok[-3:] = ["0"] * 3                # unit inertia := 0 and M at the body origin (com is already there)
        text[5] = " ".join(tok)
        with _with_env(**(dict(SBK_NOLOCAL=1) if variant == "ground" else {})):
            topo = sb.Topology(text="\n".join(text) + "\n"); bm = sb.BatchedMatter(topo, 40); bm.setPlan(plan if plan != 1 else 1)
            bm.setState(np.full((2, 40), 0.3), np.zeros((2, 40)), t=0.0)
            bm.stepBy(1e-3, 1)
            st, nbad = bm.status()
            assert nbad == 40 and np.all(st & 2), (variant, st[:4])
            bm.close(); topo.close()


@pytest.mark.parametrize("name", ["mixed7", "humanoid30", "branched_tree"])
def test_cluster_level_parallel_plan_matches_golden(name):
    """Plan 5 (a thread-block cluster per 32 instances, warps over the bodies of a level, cluster barriers between levels):
    fixed-step RKM against the reference's recorded run, the API operations (grid-level sweeps on the shared records), and
    agreement with plan 1 on a batch that is not a multiple of the warp size (several clusters, a partial last one)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = ModelInfo(str(g["text"]))
    ref = info.split_eval_out(g["eval_out"])
    got = run_eval(info, g["eval_in"], plan=5)
    for k in ref:
        assert rel_err(got[k], ref[k]) < TOL, (name, k, rel_err(got[k], ref[k]))
    y0, yref = g["step_in"], g["step_out"]
    n, ny = y0.shape[0], info.nq + info.nu
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, n); bm.setPlan(5); assert bm.getPlan() == 5
    bm.setState(soa(y0[:, :info.nq]), soa(y0[:, info.nq:]), t=0.0)
    err = bm.stepBy(float(g["h"]), int(g["nsteps"]), want_err_norm=True)
    q, u, t = bm.getState()
    assert np.all(np.isfinite(err)) and np.allclose(t, float(g["h"]) * int(g["nsteps"]))
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref[:, :ny]) < 1e-10
    bm.close()
    nb = 300
    qq, uu = info.random_states(nb, 31, q_scale=0.5)
    out = {}
    for plan in (1, 5):
        bm = sb.BatchedMatter(topo, nb); bm.setPlan(plan)
        bm.setState(soa(qq), soa(uu), t=0.0)
        e = bm.stepBy(1e-3, 3, want_err_norm=True)
        a, b, _ = bm.getState()
        out[plan] = (a, b, e)
        bm.close()
    assert rel_err(out[5][0], out[1][0]) < 1e-10 and rel_err(out[5][1], out[1][1]) < 1e-10
    assert np.allclose(out[5][2], out[1][2], rtol=1e-3, atol=1e-14)
    topo.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cluster,per_group", [(8, 1), (8, 2), (4, 4), (4, 2), (2, 1), (1, 1)])
def test_cluster_plan_schedules_agree_with_thread_per_instance(cluster, per_group):
    """Plan 5 under every shape of its schedule (cluster size, clusters per group of 32 instances: with and without the
    cross-cluster barrier, CTA-local levels, top on one CTA or on a cluster) steps a 600-body tree to the same state as the
    thread-per-instance plan; the batch has a partial last group."""
    info = ModelInfo(sb.model_text("branched_tree", 600))
    n = 72
    q, u = info.random_states(n, 77, q_scale=0.5)
    topo = sb.Topology(text=info.text)
    out = {}
    for plan in (1, 5):
        old = {k: os.environ.get(k) for k in ("SBK_CLUSTER", "SBK_CLUSTERS_PER_GROUP")}
        if plan == 5:
            os.environ["SBK_CLUSTER"] = str(cluster); os.environ["SBK_CLUSTERS_PER_GROUP"] = str(per_group)
        try:
            bm = sb.BatchedMatter(topo, n); bm.setPlan(plan); assert bm.getPlan() == plan
            bm.setState(soa(q), soa(u), t=0.0)
            e = bm.stepBy(5e-4, 3, want_err_norm=True)
            a, b, _ = bm.getState()
            st, nbad = bm.status(); assert nbad == 0
            out[plan] = (a, b, e)
            bm.close()
        finally:
            for k, v in old.items():
                if v is None: os.environ.pop(k, None)
                else: os.environ[k] = v
    assert rel_err(out[5][0], out[1][0]) < 1e-10 and rel_err(out[5][1], out[1][1]) < 1e-10
    assert np.allclose(out[5][2], out[1][2], rtol=1e-3, atol=1e-14)
    topo.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("seed,plan", [(0, 1), (1, 1), (2, 1), (3, 5), (4, 5), (5, 4), (6, 3)])
def test_random_trees_match_live_reference(seed, plan):
    """Irregular random trees of Pin / Universal / Ball bodies (1-4 children per body) through the C ABI against the real Simbody:
    every operator of `run_eval`, then five RKM steps -- thread-per-instance, cluster, grid-level and CTA-per-instance plans."""
    rng = np.random.default_rng(9000 + seed)
    nb = int(rng.integers(150, 500)) if plan == 5 else int(rng.integers(5, 70))
    info = ModelInfo(random_tree_text(rng, nb, int(rng.integers(1, 5)))); ny = info.nq + info.nu
    ein = info.random_eval_input(5, 300 + seed, q_scale=0.6)
    ref = info.split_eval_out(RefDriver().eval(info, ein))
    got = run_eval(info, ein, plan=plan)
    for k in ref:
        assert rel_err(got[k], ref[k]) < 1e-10, (seed, plan, k, rel_err(got[k], ref[k]))
    y = ein[:, :ny]
    yref = RefDriver().step(info, y, 5e-4, 5)[:, :ny]
    topo = sb.Topology(text=info.text); bm = sb.BatchedMatter(topo, y.shape[0]); bm.setPlan(plan); assert bm.getPlan() == plan
    bm.setState(soa(y[:, :info.nq]), soa(y[:, info.nq:]), t=0.0)
    bm.stepBy(5e-4, 5)
    q, u, _ = bm.getState()
    st, nbad = bm.status(); assert nbad == 0
    assert rel_err(np.concatenate([q.T, u.T], axis=1), yref) < 1e-10, (seed, plan)
    bm.close(); topo.close()
