import os, sys, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ.setdefault("FS2D_ORACLE_THREADS", "1")
import numpy as np
import helpers as H
from flipsolver2d_b200 import capi, scenes, host_api
from oracle import ref
tmp = tempfile.mkdtemp()
for name, scene, K in (("src64", scenes.source_sink(64, "flip"), 2), ("smoke64", scenes.smoke_test(64), 3)):
    scene["settings"]["density"] = scene["settings"].get("density", 1.0)
    path = os.path.join(tmp, name + ".json")
    s = H.make_ref(ref, scene, path, frames=0)
    h = host_api.Solver(path, convergence_threads=1)
    for f in range(3):
        s.step_frame(); h.step_frame()
        d = h.device(K)
        print(name, "frame", f, "ref", s.particle_count(), "dev", h.particle_count(), "substeps", s.stats()["substeps"], h.stats()["substeps"],
              "counts equal", np.array_equal(s.grid("COUNTS"), d.download("COUNTS")),
              "mat equal", np.array_equal(s.grid("MATERIAL"), d.download("MATERIAL")),
              "src cells", int((s.grid("MATERIAL") == 0x41).sum()))
        rc, dc = s.grid("COUNTS"), d.download("COUNTS")
        bad = np.nonzero(rc != dc)[0]
        if len(bad):
            print("   count diffs at", bad[:10], rc[bad[:10]], dc[bad[:10]], "mat", s.grid("MATERIAL")[bad[:10]])
    # stage-level reseed on synced state
    d2 = H.make_device(s, scene)
    pos, vel, props, _ = s.particles(); s.set_particles(pos, vel, props)
    H.sync_state(s, d2, scene["settings"]["simType"])
    s.stage("COUNT_PARTICLES"); d2.stage("count_particles")
    before = s.particle_count()
    s.stage("RESEED")
    import ctypes as C
    n = C.c_int64(0); d2.L.fs2d_reseed_plan(d2.h, C.byref(n))
    print(name, "stage reseed: ref added", s.particle_count() - before, "device plans", n.value, "counts equal", np.array_equal(s.grid("COUNTS"), d2.download("COUNTS")))

print("---- misplaced check")
scene = scenes.source_sink(64, "flip")
path = os.path.join(tmp, "mis.json")
s = H.make_ref(ref, scene, path, frames=0)
for f in range(2):
    s.step_frame()
    pos, vel, props, bins = s.particles()
    binsJ = (s.J + 2) // 3
    pb = (pos[:, 0].astype(np.int64) // 3) * binsJ + pos[:, 1].astype(np.int64) // 3
    st = s.stats()
    print("frame", f, "particles", len(pos), "misplaced", int((pb != bins).sum()), "density iters", st["density_iters"], "pressure iters", st["pressure_iters"])
