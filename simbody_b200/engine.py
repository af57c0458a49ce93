"""Host-side mirror of the reference's operator surface over the C ABI.

`BatchedMatter` keeps the method names of SimbodyMatterSubsystem / Integrator for the hot path
(realizePositionKinematics, realizeVelocityKinematics, realizeArticulatedBodyInertias,
calcAcceleration, multiplyByM, multiplyByMInv, calcResidualForceIgnoringConstraints, stepBy)
with a leading batch dimension.  Arrays are numpy float64, slot-major ("SoA"): q is [nq, N].
Everything is computed by libsbk.so on the GPU; this file only marshals pointers.
"""
import ctypes

import numpy as np

from . import capi
from .capi import SbkError, c_double_p, check, load_library


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def model_text(name, n=0):
    """Text of a built-in model (simbody_b200/host/model_spec.h): double_pendulum, pin_chain,
    humanoid30, branched_tree, mixed7."""
    lib = load_library()
    need = lib.sbk_model_text(name.encode(), n, None, 0)
    if need < 0:
        raise SbkError(1, lib.sbk_last_error().decode())
    buf = ctypes.create_string_buffer(need)
    lib.sbk_model_text(name.encode(), n, buf, need)
    return buf.value.decode()


class Topology:
    def __init__(self, text=None, bodies=None, forces=None, use_euler_angles=False):
        self.lib = load_library()
        if text is not None:
            self.handle = self.lib.sbk_topology_from_text(text.encode())
        else:
            nb, nf = len(bodies), len(forces or [])
            barr = (capi.BodyDesc * nb)(*bodies)
            farr = (capi.ForceDesc * max(nf, 1))(*(forces or []))
            self.handle = self.lib.sbk_topology_create_ex(barr, nb, farr if nf else None, nf, 1 if use_euler_angles else 0)
        if not self.handle:
            raise SbkError(3, self.lib.sbk_last_error().decode())
        v = [ctypes.c_int() for _ in range(5)]
        check(self.lib, self.lib.sbk_topology_counts(self.handle, *[ctypes.byref(x) for x in v]))
        self.nb, self.nq, self.nu, self.nquat, self.nlevels = [x.value for x in v]
        arrs = [(ctypes.c_int * self.nb)() for _ in range(5)]
        check(self.lib, self.lib.sbk_topology_slots(self.handle, *arrs))
        self.q0, self.nq_of, self.u0, self.nu_of, self.level = [list(a) for a in arrs]

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sbk_topology_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedMatter:
    """N instances of one lowered system on one GPU."""

    def __init__(self, topology, n_instances, device=0, stream=None):
        self.lib = topology.lib
        self.topo = topology
        self.N = int(n_instances)
        self.handle = self.lib.sbk_batch_create(topology.handle, self.N, device, stream)
        if not self.handle:
            raise SbkError(4, self.lib.sbk_last_error().decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sbk_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        check(self.lib, rc)

    def _vec(self, a, rows, name, optional=False):
        if a is None:
            if optional:
                return None
            raise SbkError(1, "%s is required" % name)
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (rows, self.N):
            # mirrors SimTK_APIARGCHECK on vector lengths (SimbodyMatterSubsystem.cpp:158-167)
            raise SbkError(1, "%s has shape %s, expected (%d, %d)" % (name, a.shape, rows, self.N))
        return a

    # ---- state ------------------------------------------------------------------------------
    def setState(self, q=None, u=None, t=None):
        q = self._vec(q, self.topo.nq, "q", True)
        u = self._vec(u, self.topo.nu, "u", True)
        tt = None if t is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (self.N,)))
        self._chk(self.lib.sbk_set_state(self.handle, _dp(q), _dp(u), _dp(tt)))

    def getState(self):
        q = np.empty((self.topo.nq, self.N)); u = np.empty((self.topo.nu, self.N)); t = np.empty(self.N)
        self._chk(self.lib.sbk_get_state(self.handle, _dp(q), _dp(u), _dp(t)))
        return q, u, t

    def setStateAoS(self, q, u):
        q = np.ascontiguousarray(q, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        assert q.shape == (self.N, self.topo.nq) and u.shape == (self.N, self.topo.nu)
        self._chk(self.lib.sbk_set_state_aos(self.handle, _dp(q), _dp(u)))

    def getStateAoS(self):
        q = np.empty((self.N, self.topo.nq)); u = np.empty((self.N, self.topo.nu))
        self._chk(self.lib.sbk_get_state_aos(self.handle, _dp(q), _dp(u)))
        return q, u

    def devicePointers(self):
        p = [ctypes.c_void_p() for _ in range(3)]
        self._chk(self.lib.sbk_state_device_ptrs(self.handle, *[ctypes.byref(x) for x in p]))
        return [x.value for x in p]

    def setPlan(self, plan):
        """0 auto, 1 thread-per-instance, 2 register-resident fused (tiny chains), 3 level-parallel CTA per instance,
        4 grid-level-parallel integrator (wide trees, small batches)."""
        self._chk(self.lib.sbk_batch_set_plan(self.handle, int(plan)))

    def getPlan(self):
        return int(self.lib.sbk_batch_get_plan(self.handle))

    def synchronize(self):
        self._chk(self.lib.sbk_synchronize(self.handle))

    # ---- realize (SimbodyMatterSubsystem.h:2691,2705,2730; System::realize) ---------------------
    def realizePositionKinematics(self):
        self._chk(self.lib.sbk_realize_position(self.handle))

    def realizeVelocityKinematics(self):
        self._chk(self.lib.sbk_realize_velocity(self.handle))

    def realizeArticulatedBodyInertias(self):
        self._chk(self.lib.sbk_realize_articulated_body_inertias(self.handle))

    def realizeAcceleration(self):
        self._chk(self.lib.sbk_realize_acceleration(self.handle))

    def _get(self, fn, rows):
        out = np.empty((rows, self.N))
        self._chk(fn(self.handle, _dp(out)))
        return out

    def getUDot(self):
        return self._get(self.lib.sbk_get_udot, self.topo.nu)

    def getQDot(self):
        return self._get(self.lib.sbk_get_qdot, self.topo.nq)

    def getQDotDot(self):
        return self._get(self.lib.sbk_get_qdotdot, self.topo.nq)

    def getQErr(self):
        return self._get(self.lib.sbk_get_qerr, self.topo.nquat) if self.topo.nquat else np.empty((0, self.N))

    def getBodyTransforms(self):
        return self._get(self.lib.sbk_get_body_transforms, self.topo.nb * 12).reshape(self.topo.nb, 12, self.N)

    def getBodyVelocities(self):
        return self._get(self.lib.sbk_get_body_velocities, self.topo.nb * 6).reshape(self.topo.nb, 6, self.N)

    def getBodyAccelerations(self):
        return self._get(self.lib.sbk_get_body_accelerations, self.topo.nb * 6).reshape(self.topo.nb, 6, self.N)

    def getAppliedForces(self):
        f = np.empty((self.topo.nu, self.N)); F = np.empty((self.topo.nb * 6, self.N))
        self._chk(self.lib.sbk_get_applied_forces(self.handle, _dp(f), _dp(F)))
        return f, F.reshape(self.topo.nb, 6, self.N)

    def calcEnergy(self):
        """(kinetic, potential) per instance: MultibodySystem::calcKineticEnergy / calcPotentialEnergy."""
        ke = np.empty(self.N); pe = np.empty(self.N)
        self._chk(self.lib.sbk_calc_energy(self.handle, _dp(ke), _dp(pe)))
        return ke, pe

    def calcMobilizerReactionForces(self):
        """SimbodyMatterSubsystem::calcMobilizerReactionForces (SimbodyMatterSubsystem.h:2479): [nb, 6, N] spatial
        forces at the M frame origins, in Ground; needs realizeAcceleration()."""
        F = np.empty((self.topo.nb * 6, self.N))
        self._chk(self.lib.sbk_calc_mobilizer_reaction_forces(self.handle, _dp(F)))
        return F.reshape(self.topo.nb, 6, self.N)

    def calcCompositeBodyInertias(self):
        """SimbodyMatterSubsystem::calcCompositeBodyInertias: [nb, 10, N] = mass, com(3), unit inertia xx yy zz xy xz yz
        about each body origin, in Ground; position stage."""
        R = np.empty((self.topo.nb * 10, self.N))
        self._chk(self.lib.sbk_calc_composite_body_inertias(self.handle, _dp(R)))
        return R.reshape(self.topo.nb, 10, self.N)

    def multiplyBySystemJacobian(self, v):
        """J v: [nu, N] -> [nb, 6, N] (SimbodyMatterSubsystem.h:554); position stage."""
        v = self._vec(v, self.topo.nu, "v")
        out = np.empty((self.topo.nb * 6, self.N))
        self._chk(self.lib.sbk_multiply_by_system_jacobian(self.handle, _dp(v), _dp(out)))
        return out.reshape(self.topo.nb, 6, self.N)

    def multiplyBySystemJacobianTranspose(self, F_G):
        """~J F: [nb, 6, N] -> [nu, N] (SimbodyMatterSubsystem.h:646); position stage."""
        F = self._vec(np.asarray(F_G).reshape(-1, self.N), self.topo.nb * 6, "F_G")
        out = np.empty((self.topo.nu, self.N))
        self._chk(self.lib.sbk_multiply_by_system_jacobian_transpose(self.handle, _dp(F), _dp(out)))
        return out

    # ---- operators (SimbodyMatterSubsystem.h:2141,1262,1343,2234) --------------------------------
    def calcAcceleration(self, appliedMobilityForces=None, appliedBodyForces=None):
        f = self._vec(appliedMobilityForces, self.topo.nu, "appliedMobilityForces", True)
        F = appliedBodyForces
        if F is not None:
            F = self._vec(np.asarray(F).reshape(-1, self.N), self.topo.nb * 6, "appliedBodyForces")
        udot = np.empty((self.topo.nu, self.N)); A = np.empty((self.topo.nb * 6, self.N))
        self._chk(self.lib.sbk_calc_acceleration(self.handle, _dp(f), _dp(F), _dp(udot), _dp(A)))
        return udot, A.reshape(self.topo.nb, 6, self.N)

    calcAccelerationIgnoringConstraints = calcAcceleration   # identical for tree systems (m == 0)

    def multiplyByM(self, a):
        a = self._vec(a, self.topo.nu, "a")
        out = np.empty_like(a)
        self._chk(self.lib.sbk_multiply_by_M(self.handle, _dp(a), _dp(out)))
        return out

    def multiplyByMInv(self, v):
        v = self._vec(v, self.topo.nu, "v")
        out = np.empty_like(v)
        self._chk(self.lib.sbk_multiply_by_MInv(self.handle, _dp(v), _dp(out)))
        return out

    def calcResidualForceIgnoringConstraints(self, appliedMobilityForces=None, appliedBodyForces=None, knownUdot=None):
        f = self._vec(appliedMobilityForces, self.topo.nu, "appliedMobilityForces", True)
        F = appliedBodyForces
        if F is not None:
            F = self._vec(np.asarray(F).reshape(-1, self.N), self.topo.nb * 6, "appliedBodyForces")
        ud = self._vec(knownUdot, self.topo.nu, "knownUdot", True)
        out = np.empty((self.topo.nu, self.N))
        self._chk(self.lib.sbk_calc_residual_force(self.handle, _dp(f), _dp(F), _dp(ud), _dp(out)))
        return out

    # ---- RungeKuttaMersonIntegrator, fixed step (Integrator.h:226,352) ---------------------------
    def stepBy(self, h, nsteps=1, accuracy=1e-3, constraint_tol=None, use_infinity_norm=False,
               project_every_step=False, want_err_norm=False):
        o = capi.RkmOpts(accuracy, accuracy / 10 if constraint_tol is None else constraint_tol,
                         int(use_infinity_norm), int(project_every_step))
        err = np.empty(self.N) if want_err_norm else None
        self._chk(self.lib.sbk_rkm_step(self.handle, float(h), int(nsteps), ctypes.byref(o), _dp(err)))
        return err

    def stepTo(self, t_final, accuracy=1e-3, constraint_tol=None, init_step=0.01, min_step=-1.0, max_step=-1.0,
               allow_interpolation=False, use_infinity_norm=False, project_every_step=False, max_attempts=0):
        """Integrator::stepTo with error control (every instance keeps its own step size).
        Returns (steps_taken, steps_attempted, last_step) per instance."""
        o = capi.AdaptiveOpts(accuracy, accuracy / 10 if constraint_tol is None else constraint_tol, init_step, min_step, max_step,
                              int(use_infinity_norm), int(project_every_step), int(allow_interpolation), int(max_attempts))
        steps = np.zeros(self.N, dtype=np.int32); att = np.zeros(self.N, dtype=np.int32); last = np.zeros(self.N)
        ip = ctypes.POINTER(ctypes.c_int32)
        self._chk(self.lib.sbk_rkm_adaptive(self.handle, float(t_final), ctypes.byref(o), steps.ctypes.data_as(ip),
                                            att.ctypes.data_as(ip), _dp(last)))
        return steps, att, last

    def stats(self):
        v = [ctypes.c_int64() for _ in range(3)]
        self._chk(self.lib.sbk_rkm_stats(self.handle, *[ctypes.byref(x) for x in v]))
        return {"steps_taken": v[0].value, "realizations": v[1].value, "q_projections": v[2].value}

    def status(self):
        st = np.zeros(self.N, dtype=np.int32); nbad = ctypes.c_int64()
        self._chk(self.lib.sbk_get_status(self.handle, st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.byref(nbad)))
        return st, nbad.value

    def integratorKernelName(self):
        buf = ctypes.create_string_buffer(256)
        self._chk(self.lib.sbk_integrator_kernel_name(self.handle, buf, 256))
        return buf.value.decode()

    def launchCount(self):
        return int(self.lib.sbk_launch_count(self.handle))

    def lastKernelMs(self):
        return float(self.lib.sbk_last_kernel_ms(self.handle))
