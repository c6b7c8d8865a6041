#!/usr/bin/env python3
"""Minimal driver for profilers: load a BASELINE scene through the host mirror and run substeps.

  python tools/step_probe.py [res] [steps] [sim] [--dense] [--viscous]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flipsolver2d_b200 import host_api, scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sim = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "flip"
if sim == "smoke":
    sc = scenes.smoke_test(res, ppc=4, parameter_handling="grid")
else:
    sc = scenes.dam_break(res, sim, ppc=8, pic_ratio=0.03, viscosity_enabled="--viscous" in sys.argv)
path = scenes.write_scene(sc, "/tmp/step_probe_%d_%s.json" % (res, sim))
s = host_api.Solver(path, quiet=True)
s.prepare()
d = s.device(num_properties={"flip": 2, "nbflip": 2, "smoke": 3, "fire": 4}[sim])
if "--dense" in sys.argv:
    d.pcg_set_dense(True)
for _ in range(steps):
    s.step_substep()
d.synchronize()
st = s.stats()
print("substeps", steps, "particles", s.particle_count(), "launches", s.kernel_launches(),
      {n: round(float(st["timings"][k]) / max(st["substeps"], 1), 3) for k, n in enumerate(host_api.STAGES)})
