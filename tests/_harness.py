"""Shared test plumbing: model texts, the reference driver (oracle/_ref) and I/O layouts.

Nothing here is product code.  `RefDriver` executes oracle/_ref/ref_driver, the harness linked
against the UNMODIFIED reference (simbody 3.9.0) compiled by oracle/Makefile.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")

JOINT_NQ = {"PIN": 1, "SLIDER": 1, "UNIVERSAL": 2, "BALL": 4, "FREE": 7}
JOINT_NU = {"PIN": 1, "SLIDER": 1, "UNIVERSAL": 2, "BALL": 3, "FREE": 6}


def have_ref():
    return os.path.exists(REF_DRIVER) and os.path.exists(os.path.join(REF_DIR, "libsimbody_ref.so"))


class ModelInfo:
    """Parsed view of a model text (format of simbody_b200/host/model_spec.h)."""

    def __init__(self, text):
        self.text = text
        self.joints, self.parents = [], []
        for line in text.splitlines():
            tok = line.split()
            if tok and tok[0] == "body":
                self.parents.append(int(tok[2]))
                self.joints.append(tok[3])
        self.nb = len(self.joints)
        self.q0, self.u0 = [], []
        nq = nu = 0
        for j in self.joints:
            self.q0.append(nq)
            self.u0.append(nu)
            nq += JOINT_NQ.get(j, 0)
            nu += JOINT_NU.get(j, 0)
        self.nq, self.nu = nq, nu
        self.nquat = sum(1 for j in self.joints if j in ("BALL", "FREE"))
        self.quat_q0 = [self.q0[i] for i, j in enumerate(self.joints) if j in ("BALL", "FREE")]

    # ---- layouts of `ref_driver eval` (oracle/ref_driver.cpp) --------------------------------
    @property
    def eval_in_stride(self):
        return self.nq + 5 * self.nu + 6 * self.nb

    def eval_out_fields(self):
        nb, nq, nu, nquat = self.nb, self.nq, self.nu, self.nquat
        return [("qdot", nq), ("udot", nu), ("qdotdot", nq), ("qerr", nquat), ("X_GB", nb * 12),
                ("V_GB", nb * 6), ("A_GB", nb * 6), ("fmob_sys", nu), ("Fbody_sys", nb * 6),
                ("Ma", nu), ("MInvv", nu), ("resid", nu), ("resid0", nu), ("udot_op", nu),
                ("A_GB_op", nb * 6)]

    @property
    def eval_out_stride(self):
        return sum(n for _, n in self.eval_out_fields())

    def split_eval_out(self, out):
        out = np.asarray(out).reshape(-1, self.eval_out_stride)
        res, o = {}, 0
        for name, n in self.eval_out_fields():
            res[name] = out[:, o:o + n]
            o += n
        return res

    def random_states(self, n, seed, q_scale=1.0, u_scale=1.0):
        """Seeded q,u with unit quaternions; returns (q [n,nq], u [n,nu])."""
        rng = np.random.default_rng(seed)
        q = rng.uniform(-1, 1, size=(n, self.nq)) * q_scale
        u = rng.uniform(-1, 1, size=(n, self.nu)) * u_scale
        for s in self.quat_q0:
            quat = rng.normal(size=(n, 4))
            q[:, s:s + 4] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
        return q, u

    def random_eval_input(self, n, seed, **kw):
        q, u = self.random_states(n, seed, **kw)
        rng = np.random.default_rng(seed + 7919)
        rest = rng.uniform(-1, 1, size=(n, 4 * self.nu + 6 * self.nb))
        return np.ascontiguousarray(np.concatenate([q, u, rest], axis=1))


class RefDriver:
    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/ref_driver missing: run `make -C oracle` where /root/reference exists")
        self.env = dict(os.environ)
        self.env["LD_LIBRARY_PATH"] = REF_DIR + ":" + self.env.get("LD_LIBRARY_PATH", "")

    def _run(self, args, **kw):
        r = subprocess.run([REF_DRIVER] + [str(a) for a in args], env=self.env, capture_output=True, text=True, **kw)
        if r.returncode != 0:
            raise RuntimeError("ref_driver %s failed (%d): %s" % (args[0], r.returncode, r.stderr[-2000:]))
        return r.stdout

    def _with_model(self, text, fn):
        with tempfile.TemporaryDirectory() as d:
            mp = os.path.join(d, "model.txt")
            with open(mp, "w") as f:
                f.write(text)
            return fn(d, mp)

    def lower(self, text):
        return self._with_model(text, lambda d, mp: self._run(["lower", mp]))

    def slots(self, text):
        return self._with_model(text, lambda d, mp: self._run(["slots", mp]))

    def eval(self, info, inp):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        n = inp.shape[0]

        def go(d, mp):
            ip, op = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            inp.tofile(ip)
            self._run(["eval", mp, ip, op, n])
            return np.fromfile(op, dtype=np.float64).reshape(n, info.eval_out_stride)
        return self._with_model(info.text, go)

    def energy(self, info, y):
        """-> [n, 2]: kinetic, potential energy (MultibodySystem::calcKineticEnergy / calcPotentialEnergy)"""
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]

        def go(d, mp):
            ip, op = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            y.tofile(ip)
            self._run(["energy", mp, ip, op, n])
            return np.fromfile(op, dtype=np.float64).reshape(n, 2)
        return self._with_model(info.text, go)

    def step(self, info, y, h, nsteps, accuracy=-1):
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]

        def go(d, mp):
            ip, op = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            y.tofile(ip)
            self._run(["step", mp, ip, op, n, repr(float(h)), nsteps, repr(float(accuracy))])
            return np.fromfile(op, dtype=np.float64).reshape(n, info.nq + info.nu + 3)
        return self._with_model(info.text, go)

    def adaptive(self, info, y, t_final, accuracy=-1, allow_interpolation=True):
        """-> [n, ny+5]: advanced q,u | steps taken | attempted | realizations | last step | advanced time"""
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]

        def go(d, mp):
            ip, op = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            y.tofile(ip)
            self._run(["adaptive", mp, ip, op, n, repr(float(t_final)), repr(float(accuracy)), int(allow_interpolation)])
            return np.fromfile(op, dtype=np.float64).reshape(n, info.nq + info.nu + 5)
        return self._with_model(info.text, go)

    def bench(self, info, y, h, nsteps, threads):
        import json
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]

        def go(d, mp):
            ip = os.path.join(d, "in.bin")
            y.tofile(ip)
            return json.loads(self._run(["bench", mp, ip, n, repr(float(h)), nsteps, threads]))
        return self._with_model(info.text, go)


class HostEmu:
    """ctypes view of build/libsbk_hostemu.so (tests/hostemu/hostemu.cpp) -- test-only."""

    def __init__(self):
        path = os.path.join(ROOT, "build", "libsbk_hostemu.so")
        self.lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        self.lib.emu_eval.argtypes = [ctypes.c_char_p, ctypes.c_int, dp, dp]
        self.lib.emu_step.argtypes = [ctypes.c_char_p, ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        self.lib.emu_adaptive.argtypes = [ctypes.c_char_p, ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_int, ctypes.c_int]
        self.lib.emu_model_text.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]

    def model_text(self, name, n=0):
        need = self.lib.emu_model_text(name.encode(), n, None, 0)
        assert need > 0
        buf = ctypes.create_string_buffer(need)
        self.lib.emu_model_text(name.encode(), n, buf, need)
        return buf.value.decode()

    def eval(self, info, inp):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        n = inp.shape[0]
        out = np.zeros((n, info.eval_out_stride))
        dp = ctypes.POINTER(ctypes.c_double)
        rc = self.lib.emu_eval(info.text.encode(), n, inp.ctypes.data_as(dp), out.ctypes.data_as(dp))
        assert rc == 0, rc
        return out

    def step(self, info, y, h, nsteps, accuracy=1e-3, cons_tol=None, inf_norm=0, project_every=0, lean=1):
        """lean=1 runs the integrator's LEAN body wrappers (carry links), lean=0 the full-record ones."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]
        out = np.zeros((n, info.nq + info.nu + 2))
        dp = ctypes.POINTER(ctypes.c_double)
        rc = self.lib.emu_step(info.text.encode(), n, y.ctypes.data_as(dp), out.ctypes.data_as(dp), h, nsteps,
                               accuracy, accuracy / 10 if cons_tol is None else cons_tol, inf_norm, project_every, lean)
        assert rc == 0, rc
        return out


def _emu_adaptive(self, info, y, t_final, accuracy=1e-3, init_step=0.01, allow_interpolation=True, fused=False):
    """-> [n, ny+5] in the layout of RefDriver.adaptive (realizations = steps + 4*attempts)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    out = np.zeros((n, info.nq + info.nu + 5))
    dp = ctypes.POINTER(ctypes.c_double)
    rc = self.lib.emu_adaptive(info.text.encode(), n, y.ctypes.data_as(dp), out.ctypes.data_as(dp), t_final, accuracy, init_step,
                               int(allow_interpolation), int(fused))
    assert rc == 0, rc
    return out


HostEmu.adaptive = _emu_adaptive


def rel_err(a, b):
    """max |a-b| / max(1, max|b|) -- the relative measure used for the 1e-10 parity bar."""
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


class COracle:
    """ctypes view of build/libsbk_oracle.so (oracle/sbk_oracle.c, the plain-C restatement)."""

    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(ROOT, "build", "libsbk_oracle.so"))

    @staticmethod
    def _arrays(text):
        jt = {"GROUND": 0, "PIN": 1, "SLIDER": 2, "UNIVERSAL": 3, "BALL": 4, "FREE": 5}
        parent, joint, mass, com, ui, xpf, xbm = [], [], [], [], [], [], []
        fk, fb, fc, fa, fbb, fd = [], [], [], [], [], []
        for line in text.splitlines():
            t = line.split()
            if not t:
                continue
            if t[0] == "body":
                v = [float(x) for x in t[4:]]
                parent.append(int(t[2])); joint.append(jt[t[3]]); mass.append(v[0]); com += v[1:4]; ui += v[4:10]
                xpf += v[10:22]; xbm += v[22:34]
            elif t[0] == "gravity":
                fk.append(1); fb.append(-1); fc.append(0); fa.append(float(t[1])); fbb.append(0.0); fd += [float(x) for x in t[2:5]]
            elif t[0] == "spring":
                fk.append(2); fb.append(int(t[1])); fc.append(int(t[2])); fa.append(float(t[3])); fbb.append(float(t[4])); fd += [0.0] * 3
            elif t[0] == "damper":
                fk.append(3); fb.append(int(t[1])); fc.append(int(t[2])); fa.append(float(t[3])); fbb.append(0.0); fd += [0.0] * 3
        i32 = lambda a: np.ascontiguousarray(a if len(a) else [0], dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a if len(a) else [0.0], dtype=np.float64)
        return (len(parent), i32(parent), i32(joint), f64(mass), f64(com), f64(ui), f64(xpf), f64(xbm),
                len(fk), i32(fk), i32(fb), i32(fc), f64(fa), f64(fbb), f64(fd))

    @staticmethod
    def _cargs(arrs):
        out = []
        for a in arrs:
            if isinstance(a, np.ndarray):
                out.append(a.ctypes.data_as(ctypes.c_void_p))
            else:
                out.append(ctypes.c_int(a))
        return out

    def eval(self, info, inp):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        n = inp.shape[0]
        out = np.zeros((n, info.eval_out_stride))
        arrs = self._arrays(info.text)
        rc = self.lib.oracle_eval(*self._cargs(arrs), ctypes.c_int(n), inp.ctypes.data_as(ctypes.c_void_p),
                                  out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        return out

    def step(self, info, y, h, nsteps, accuracy=1e-3, cons_tol=None, inf_norm=0, project_every=0):
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]
        out = np.zeros((n, info.nq + info.nu + 2))
        arrs = self._arrays(info.text)
        rc = self.lib.oracle_step(*self._cargs(arrs), ctypes.c_int(n), y.ctypes.data_as(ctypes.c_void_p),
                                  out.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(h), ctypes.c_int(nsteps),
                                  ctypes.c_double(accuracy), ctypes.c_double(accuracy / 10 if cons_tol is None else cons_tol),
                                  ctypes.c_int(inf_norm), ctypes.c_int(project_every))
        assert rc == 0
        return out
