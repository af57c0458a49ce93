"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver).

Run where /root/reference was compiled by `make -C oracle`:   python tests/golden/make_golden.py
Each fixture holds the model text, seeded inputs, and the reference's outputs for
  * realize(Acceleration) + the matter operators   (ref_driver eval)
  * fixed-step RungeKuttaMerson                      (ref_driver step)
  * calcMobilizerReactionForces, multiplyBySystemJacobian[Transpose]   (ref_driver extras)
so that the parity tests can run on machines without the reference sources.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _harness import HostEmu, ModelInfo, RefDriver  # noqa: E402

CASES = [  # name, size param, n eval states, (h, nsteps), q_scale
    ("double_pendulum", 0, 8, (1e-3, 25), 3.0),
    ("pin_chain", 50, 4, (1e-3, 5), 1.0),
    ("mixed7", 0, 8, (1e-3, 25), 1.0),
    ("mixed7e", 0, 6, (1e-3, 25), 0.6),          # SimbodyMatterSubsystem::setUseEulerAngles: Ball / Free with x-y-z angles
    ("ugdamp5", 0, 6, (1e-3, 20), 0.7),          # Force::UniformGravity + Force::GlobalDamper
    ("welded8", 0, 6, (1e-3, 20), 0.7),          # MobilizedBody::Weld inside the chain and as a leaf
    ("cartesian8", 0, 6, (1e-3, 20), 0.7),       # MobilizedBody::Planar / Cylinder / Translation
    ("twopoint7", 0, 6, (1e-3, 20), 0.7),        # Force::TwoPointLinearSpring / TwoPointLinearDamper between bodies and to Ground
    ("humanoid30", 0, 4, (1e-3, 10), 0.5),
    ("branched_tree", 100, 2, (5e-4, 4), 0.5),
    ("branched_tree", 1000, 2, (5e-4, 2), 0.5),  # BASELINE config 5 at full size (11 levels, width 489): file branched_tree1000.npz
]


def main():
    emu, ref = HostEmu(), RefDriver()
    only = sys.argv[1:]                       # optional: fixture file stems to (re)generate
    for name, n, neval, (h, nsteps), qs in CASES:
        stem = name + (str(n) if (name, n) == ("branched_tree", 1000) else "")
        if only and stem not in only:
            continue
        text = emu.model_text(name, n)
        info = ModelInfo(text)
        assert ref.lower(text) == text, "lowering must reproduce the spec for " + name
        ein = info.random_eval_input(neval, 1000 + len(name), q_scale=qs)
        eout = ref.eval(info, ein)
        q, u = info.random_states(neval, 2000 + len(name), q_scale=qs)
        y0 = np.concatenate([q, u], axis=1)
        yout = ref.step(info, y0, h, nsteps)
        path = os.path.join(HERE, "%s.npz" % stem)
        energy = ref.energy(info, ein[:, :info.nq + info.nu])      # [n, 2] kinetic, potential at the eval states
        xin = info.random_extras_input(neval, 3000 + len(name), q_scale=qs)    # reactions, J v, ~J F  (ref_driver extras)
        xout = ref.extras(info, xin)
        np.savez_compressed(path, text=np.array(text), eval_in=ein, eval_out=eout, step_in=y0, step_out=yout,
                            h=h, nsteps=nsteps, slots=np.array(ref.slots(text)), energy=energy, extras_in=xin, extras_out=xout)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
