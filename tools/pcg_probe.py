"""Quick PCG timing probe (not the bench): synthetic dam-break material grid, random rhs."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from flipsolver2d_b200 import capi

def material(res):
    m = np.full((res, res), capi.EMPTY, np.int8)
    w = int(round(res * 3 / 50))
    m[:w, :] = capi.SOLID; m[-w:, :] = capi.SOLID; m[:, :w] = capi.SOLID; m[:, -w:] = capi.SOLID
    m[res // 2: res - w, w: int(res * 13 / 50)] = capi.FLUID
    return m

for res in [int(a) for a in sys.argv[1:]] or [1024, 4096]:
    d = capi.Device(res, res, dx=50.0 / res, fluid_density=0.5, pcg_iter_limit=200)
    m = material(res)
    d.upload("MATERIAL", m)
    d.set_step_dt(1 / 30.0)
    d.stage("build_matrix")
    rng = np.random.default_rng(0)
    rhs = np.where(m.ravel() == capi.FLUID, rng.standard_normal(res * res), 0.0)
    d.upload("RHS", rhs)
    for conv in (0,):
        d.pcg_solve_device(200, 0.0); d.synchronize()
        t = time.perf_counter()
        d.pcg_solve_device(200, 0.0); d.synchronize()
        dt = time.perf_counter() - t
        n = res * res
        print("res %d: %.3f ms/iter, %.1f GB/s at 83 B/cell (10 fp64 passes + 3 B), %.1f GB/s at 98 B/cell"
              % (res, dt / 200 * 1e3, 83 * n * 200 / dt / 1e9, 98 * n * 200 / dt / 1e9))
    d.close()
