#!/usr/bin/env python3
"""Builds tests/golden/viscosity_systems.npz: inputs and outputs of the reference's own implicit-viscosity stage
(FlipSolver::applyViscosity -> Light / HeavyViscosityModel::apply, viscositymodel.cpp) on a 48x48 dam-break state,
produced by the oracle (oracle/_ref = the unmodified reference sources + the Eigen stand-in). The GPU box has no
/root/reference: the -m gpu tests replay these inputs through the CUDA stage and compare with the stored outputs, and
a CPU test pins the numpy / scipy restatements of both systems on the same vectors. Run where oracle/_ref is built.

Eigen 3.4.0 itself is not in the reference tree: these vectors pin the path against the reference's
assembly + the restated conjugate_gradient, not against Eigen's binary.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from flipsolver2d_b200 import scenes  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "viscosity_systems.npz")


def capture(heavy, res=48, frames=2):
    sc = scenes.dam_break(res, "flip", viscosity_enabled=True, fluid_viscosity=10)
    sc["settings"]["density"] = 0.02
    if heavy:
        sc["settings"]["heavyViscosity"] = True
    path = scenes.write_scene(sc, "/tmp/golden_visc_%d.json" % int(heavy))
    s = ref.RefSolver(path, strict=True)
    for _ in range(frames):
        s.step_frame()
    s.set_step_dt(1.0 / 120.0)
    p = s.params()
    out = dict(I=s.I, J=s.J, dt=np.float32(p["stepDt"]), dx=p["dx"], density=p["fluidDensity"], material=s.grid("MATERIAL"),
               viscosity=s.grid("VISCOSITY"), u0=s.grid("U"), v0=s.grid("V"))
    s.stage("VISCOSITY")
    out["u1"], out["v1"] = s.grid("U"), s.grid("V")
    s.close()
    return out


def main():
    ref.load(strict=True, threads=1)
    data = {}
    for name, heavy in (("light", False), ("heavy", True)):
        for k, v in capture(heavy).items():
            data["%s_%s" % (name, k)] = v
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
