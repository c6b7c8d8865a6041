"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_ref/libfs2d_ref*.so (oracle/ref_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product package (flipsolver2d_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

STAGE = dict(
    ADVECT=0, BUILD_MATRIX=1, PRUNE_REBIN=2, DENSITY_CORRECTION=3, P2G=4, UPDATE_SDF=5,
    UPDATE_MATERIALS=6, AFTER_TRANSFER=7, EXTRAPOLATE_SDF_IN=8, EXTRAPOLATE_VEL=9,
    SAVE_VELOCITY=10, BODY_FORCES=11, PROJECT=12, VELOCITY_FROM_SOLIDS=13, VISCOSITY=14,
    PARTICLE_UPDATE=15, COUNT_PARTICLES=16, RESEED=17, GRID_UPDATE=18, FULL_STEP=19,
    UPDATE_DENSITY_GRID=20, EXTRAPOLATE_SDF_OUT=21, FIRST_FRAME_INIT=22,
)

GRID = dict(
    U=(0, np.float32), V=(1, np.float32), U_VALID=(2, np.uint8), V_VALID=(3, np.uint8),
    SAVED_U=(4, np.float32), SAVED_V=(5, np.float32), MATERIAL=(6, np.int8),
    FLUID_SDF=(7, np.float32), SOLID_SDF=(8, np.float32), VISCOSITY=(9, np.float32),
    DENSITY=(10, np.float32), COUNTS=(11, np.int32), EMITTER_ID=(12, np.int32),
    SOLID_ID=(13, np.int32), DIVERGENCE_CONTROL=(14, np.float32), TEST=(15, np.float32),
    KNOWN_CENTERED=(16, np.uint8), TEMPERATURE=(17, np.float32), CONCENTRATION=(18, np.float32),
    FUEL=(19, np.float32),
)

_libs = {}


def lib_path(strict=True):
    return os.path.join(HERE, "_ref", "libfs2d_ref_strict.so" if strict else "libfs2d_ref.so")


def available(strict=True):
    return os.path.exists(lib_path(strict))


def load(strict=True, threads=None):
    """Load one variant. `threads` pins the reference ThreadPool size (first load wins:
    the pool is a process-wide singleton, threading/threadpool.cpp:30-34)."""
    key = bool(strict)
    if key in _libs:
        return _libs[key]
    if threads is not None:
        os.environ["FS2D_ORACLE_THREADS"] = str(int(threads))
    L = C.CDLL(lib_path(strict))
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    sig = {
        "ref_load_scene": (vp, [C.c_char_p]),
        "ref_destroy": (None, [vp]),
        "ref_thread_count": (i32, []),
        "ref_set_quiet": (None, [i32]),
        "ref_size_i": (i32, [vp]),
        "ref_size_j": (i32, [vp]),
        "ref_sim_type": (i32, [vp]),
        "ref_particle_count": (i64, [vp]),
        "ref_property_count": (i32, [vp]),
        "ref_frame_number": (i32, [vp]),
        "ref_get_params": (None, [vp, vp]),
        "ref_step_frame": (None, [vp]),
        "ref_get_stats": (None, [vp, vp, vp]),
        "ref_set_step_dt": (None, [vp, f32]),
        "ref_max_particle_velocity": (f32, [vp]),
        "ref_run_stage": (None, [vp, i32]),
        "ref_bump_frame_number": (None, [vp]),
        "ref_get_particles": (None, [vp, vp, vp, vp, vp]),
        "ref_set_particles": (None, [vp, i64, vp, vp, vp]),
        "ref_grid_size": (i64, [vp, i32]),
        "ref_get_grid": (i32, [vp, i32, vp]),
        "ref_set_grid": (i32, [vp, i32, vp]),
        "ref_get_matrix": (f64, [vp, vp, vp, vp, vp]),
        "ref_spmv": (None, [vp, vp, vp]),
        "ref_precond_apply": (None, [vp, vp, vp]),
        "ref_pcg_solve": (i32, [vp, vp, vp, i32, f64]),
        "ref_pressure_rhs": (None, [vp, vp]),
        "ref_density_rhs": (None, [vp, vp]),
        "ref_apply_pressure": (None, [vp, vp]),
        "ref_vops_dot": (f64, [vp, vp, i64]),
        "ref_vops_max_abs": (f64, [vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _libs[key] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefSolver:
    """One reference solver instance (any of the four sim types)."""

    def __init__(self, json_path, strict=True, threads=None, quiet=True):
        self.L = load(strict, threads)
        self.L.ref_set_quiet(1 if quiet else 0)
        self.h = self.L.ref_load_scene(str(json_path).encode())
        if not self.h:
            raise RuntimeError("reference failed to load scene %s" % json_path)
        self.I = self.L.ref_size_i(self.h)
        self.J = self.L.ref_size_j(self.h)
        self.N = self.I * self.J

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return self.L.ref_thread_count()

    def params(self):
        out = np.zeros(16, np.float64)
        self.L.ref_get_params(self.h, _p(out))
        keys = ["stepDt", "frameDt", "dx", "fluidDensity", "ppc", "gx", "gy", "picRatio", "cfl",
                "particleScale", "pcgIterLimit", "projectTolerance", "maxSubsteps",
                "viscosityEnabled", "parameterHandling", "fps"]
        return dict(zip(keys, out.tolist()))

    def sim_type(self):
        return self.L.ref_sim_type(self.h)

    def particle_count(self):
        return int(self.L.ref_particle_count(self.h))

    def property_count(self):
        return int(self.L.ref_property_count(self.h))

    def frame_number(self):
        return int(self.L.ref_frame_number(self.h))

    def step_frame(self):
        self.L.ref_step_frame(self.h)

    def stats(self):
        t = np.zeros(12, np.float32)
        m = np.zeros(5, np.float32)
        self.L.ref_get_stats(self.h, _p(t), _p(m))
        return dict(timings=t, frame_ms=float(m[0]), substeps=int(m[1]), pressure_iters=int(m[2]),
                    density_iters=int(m[3]), viscosity_iters=int(m[4]))

    def set_step_dt(self, dt):
        self.L.ref_set_step_dt(self.h, float(dt))

    def max_particle_velocity(self):
        return float(self.L.ref_max_particle_velocity(self.h))

    def stage(self, name):
        self.L.ref_run_stage(self.h, STAGE[name])

    def bump_frame(self):
        self.L.ref_bump_frame_number(self.h)

    def particles(self):
        n = self.particle_count()
        k = self.property_count()
        pos = np.zeros((n, 2), np.float32)
        vel = np.zeros((n, 2), np.float32)
        props = np.zeros((k, n), np.float32)
        bins = np.zeros(n, np.int32)
        self.L.ref_get_particles(self.h, _p(pos), _p(vel), _p(props), _p(bins))
        return pos, vel, props, bins

    def set_particles(self, pos, vel=None, props=None):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        props = None if props is None else np.ascontiguousarray(props, np.float32)
        self.L.ref_set_particles(self.h, n, _p(pos), _p(vel), _p(props))

    def grid(self, name):
        gid, dt = GRID[name]
        n = int(self.L.ref_grid_size(self.h, gid))
        out = np.zeros(n, dt)
        if n and self.L.ref_get_grid(self.h, gid, _p(out)) != 0:
            raise KeyError(name)
        return out

    def set_grid(self, name, data):
        gid, dt = GRID[name]
        a = np.ascontiguousarray(data, dt).ravel()
        assert a.size == int(self.L.ref_grid_size(self.h, gid)), name
        if self.L.ref_set_grid(self.h, gid, _p(a)) != 0:
            raise KeyError(name)

    def matrix(self):
        n = self.N
        is_unit = np.zeros(n, np.uint8)
        mask = np.zeros(n, np.uint8)
        count = np.zeros(n, np.uint8)
        coef = np.zeros((4, n), np.float64)
        scale = self.L.ref_get_matrix(self.h, _p(is_unit), _p(mask), _p(count), _p(coef))
        return dict(scale=scale, is_unit=is_unit, mask=mask, count=count, coef=coef)

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        self.L.ref_spmv(self.h, _p(x), _p(y))
        return y

    def precond(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        self.L.ref_precond_apply(self.h, _p(x), _p(y))
        return y

    def pcg(self, rhs, iter_limit, tol):
        rhs = np.ascontiguousarray(rhs, np.float64)
        x = np.zeros_like(rhs)
        it = self.L.ref_pcg_solve(self.h, _p(rhs), _p(x), int(iter_limit), float(tol))
        return x, int(it)

    def pressure_rhs(self):
        out = np.zeros(self.N, np.float64)
        self.L.ref_pressure_rhs(self.h, _p(out))
        return out

    def density_rhs(self):
        out = np.zeros(self.N, np.float64)
        self.L.ref_density_rhs(self.h, _p(out))
        return out

    def apply_pressure(self, p):
        p = np.ascontiguousarray(p, np.float64)
        self.L.ref_apply_pressure(self.h, _p(p))
