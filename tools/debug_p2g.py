import os, sys, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ.setdefault("FS2D_ORACLE_THREADS", "8")
import numpy as np
import helpers as H
from flipsolver2d_b200 import capi, scenes
from oracle import ref
tmp = tempfile.mkdtemp()
scene = scenes.dam_break(96, "flip")
s = H.make_ref(ref, scene, os.path.join(tmp, "a.json"), frames=2)
s.set_step_dt(1/120.)
d = H.make_device(s, scene)
pos, vel, props, bins = s.particles()
s.set_particles(pos, vel, props)
H.sync_state(s, d)
s.stage("P2G"); d.stage("particle_to_grid")
for g in ("U_VALID", "V_VALID", "KNOWN_CENTERED"):
    a, b = s.grid(g), d.download(g)
    bad = np.nonzero(a != b)[0]
    print(g, "mismatch", len(bad), bad[:20], "ref", a[bad[:20]], "dev", b[bad[:20]])
    if g == "V_VALID" and len(bad):
        J1 = s.J + 1
        print(" (i,j):", [(int(x // J1), int(x % J1)) for x in bad[:20]])
# thread ranges: 8 threads * 16 jobs
N = s.N
print("N", N, "I,J", s.I, s.J)
