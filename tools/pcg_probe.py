"""Quick PCG timing probe (not the bench): synthetic dam-break material grid, random rhs."""
import os, sys, time
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
    if os.environ.get("FS2D_PROBE_SLAB1"):
        d.slab_configure(0, 1)   # one rank in slab mode: the <MG = true> kernels without a neighbour
    d.upload("MATERIAL", m)
    d.set_step_dt(1 / 30.0)
    d.stage("build_matrix")
    rng = np.random.default_rng(0)
    rhs = np.where(m.ravel() == capi.FLUID, rng.standard_normal(res * res), 0.0)
    d.upload("RHS", rhs)
    xs = {}
    for mode in [int(m) for m in os.environ.get('FS2D_PROBE_MODES', '1,2,0').split(',')]:  # 1 resident (+ paged tiles), 2 resident without paging, 0 streaming
        d.pcg_set_resident(mode)
        d.pcg_solve_device(200, 0.0); d.synchronize()
        d.pcg_profile(True)
        for _ in range(int(os.environ.get('FS2D_PROBE_SOLVES', '5'))):
            d.pcg_solve_device(200, 0.0)
        ms, solves = d.pcg_profile_solves()
        d.pcg_profile(False)
        xs[mode] = d.download("PRESSURE").astype(np.float64).ravel()
        n = d.pcg_active_cells()
        dt = ms / solves * 1e-3
        if int(os.environ.get("FS2D_MG_DEBUG", "0")) & 8 and d.pcg_last_kernel() > 0:
            import ctypes as C
            L = capi.lib()
            L.fs2d_debug_mg_timeline.argtypes = [C.c_void_p, C.c_void_p]
            tl = np.zeros((1024, 8), np.uint64)
            L.fs2d_debug_mg_timeline(d.h, tl.ctypes.data_as(C.c_void_p))
            t = tl[101:301].astype(np.int64)
            for name, sel in (("k1", t[0::2]), ("k2", t[1::2])):
                print("   %s (CTA 0): resident tiles %.2f us, paged tiles %.2f us, barrier %.2f us" % (
                    name, np.mean(sel[:, 1] - sel[:, 0]) / 1e3, np.mean(sel[:, 2] - sel[:, 1]) / 1e3, np.mean(sel[:, 3] - sel[:, 2]) / 1e3))
        print("res %d mode %d: kernel %d, %d active tiles, %.2f us/iteration, %.1f GB/s at 83 B per active cell, x vs mode 1: %.2e"
              % (res, mode, d.pcg_last_kernel(), n // 2048, dt / 200 * 1e6, 83 * n * 200 / dt / 1e9,
                 np.linalg.norm(xs[mode] - xs[1]) / max(np.linalg.norm(xs[1]), 1e-300)), flush=True)
    d.close()
