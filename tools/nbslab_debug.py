#!/usr/bin/env python3
"""Debug aid: nbflip on one handle vs two slab ranks sharing one GPU, substep by substep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from flipsolver2d_b200 import capi, host_api, scenes

res, world, steps = 128, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 6
sc = scenes.dam_break(res, "nbflip", viscosity_enabled="--viscous" in sys.argv)
sc["solver"]["objects"][-1]["verts"] = [[15, 3], [15, 13], [40, 13], [40, 3]]
path = scenes.write_scene(sc, "/tmp/nbslab_debug.json")
single = host_api.Solver(path, quiet=True)
solvers = [host_api.Solver(path, quiet=True, slab=(r, world, world)) for r in range(world)]
host_api.connect_slabs(solvers)
single.prepare()
capi.run_ranks([lambda s=s: s.prepare() for s in solvers])
d1 = single.device(2)
devs = [s.device(2) for s in solvers]
J = single.J
for k in range(steps):
    single.step_substep()
    capi.run_ranks([lambda s=s: s.step_substep() for s in solvers])
    n1 = single.particle_count()
    ns = [s.particle_count() for s in solvers]
    line = "step %d particles single %d slabs %s (sum %d)" % (k, n1, ns, sum(ns))
    for name, per in (("MATERIAL", J), ("COUNTS", J), ("U", J), ("V", J + 1), ("VISCOSITY", J)):
        ref = d1.download(name).reshape(-1, per)
        worst = 0.0
        for rk, d in enumerate(devs):
            lo, hi, _ = d.slab_rows()
            got = d.download(name).reshape(-1, per)[lo:hi]
            diff = np.abs(got.astype(np.float64) - ref[lo:hi].astype(np.float64))
            worst = max(worst, float(diff.max()))
            if diff.max() > 0:
                rows = np.nonzero(diff.max(axis=1))[0] + lo
                line += " | %s r%d maxdiff %.3g rows %d..%d (slab %d..%d)" % (name, rk, diff.max(), rows.min(), rows.max(), lo, hi)
    print(line, flush=True)
    # band level set of the solver itself (device pointer view is not available from python: compare owned rows of a raw download)
