"""ctypes binding of include/fs2d.h (libfs2d_cuda.so).

This is plumbing for tests and bench.py: every call goes straight through the C ABI.
There is no CPU fallback -- if the CUDA library is missing or no device is present the
import / fs2d_create fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfs2d_cuda.so")

OK = 0
ERRORS = {-1: "CUDA error", -2: "bad argument", -3: "bad state", -4: "no CUDA device", -5: "communication error"}

SIM_LIQUID, SIM_SMOKE, SIM_FIRE, SIM_NBFLIP = 0, 1, 2, 3
PARAMS_PARTICLE, PARAMS_HYBRID, PARAMS_GRID = 0, 1, 2
FLUID, SOURCE, SOLID, SINK, EMPTY = 0x40, 0x41, 0x20, 0x12, 0x10

GRID = dict(
    U=(0, np.float32), V=(1, np.float32), U_VALID=(2, np.uint8), V_VALID=(3, np.uint8),
    SAVED_U=(4, np.float32), SAVED_V=(5, np.float32), MATERIAL=(6, np.int8),
    FLUID_SDF=(7, np.float32), SOLID_SDF=(8, np.float32), VISCOSITY=(9, np.float32),
    DENSITY=(10, np.float32), COUNTS=(11, np.int32), EMITTER_ID=(12, np.int32),
    SOLID_ID=(13, np.int32), DIVERGENCE_CONTROL=(14, np.float32), TEST=(15, np.float32),
    KNOWN_CENTERED=(16, np.uint8), TEMPERATURE=(17, np.float32), CONCENTRATION=(18, np.float32),
    FUEL=(19, np.float32), PRESSURE=(20, np.float64), RHS=(21, np.float64),
    SOURCE_SDF=(22, np.float32), SOURCE_SDF_ID=(23, np.int32), ADVECTED_U=(24, np.float32),
    ADVECTED_V=(25, np.float32), ADVECTED_SDF=(26, np.float32), ADVECTED_VISCOSITY=(27, np.float32),
)


class Params(C.Structure):
    _fields_ = [
        ("size_i", C.c_int32), ("size_j", C.c_int32), ("num_properties", C.c_int32),
        ("particles_per_cell", C.c_int32), ("pcg_iter_limit", C.c_int32), ("sim_type", C.c_int32),
        ("parameter_handling", C.c_int32), ("viscosity_enabled", C.c_int32),
        ("convergence_threads", C.c_int32), ("device", C.c_int32), ("viscosity_property", C.c_int32),
        ("temperature_property", C.c_int32), ("concentration_property", C.c_int32),
        ("fuel_property", C.c_int32), ("test_property", C.c_int32), ("heavy_viscosity", C.c_int32),
        ("dx", C.c_double), ("fluid_density", C.c_double), ("project_tolerance", C.c_double),
        ("gravity_x", C.c_float), ("gravity_y", C.c_float), ("pic_ratio", C.c_float),
        ("particle_scale", C.c_float), ("ambient_temperature", C.c_float),
        ("temperature_decay", C.c_float), ("concentration_decay", C.c_float),
        ("buoyancy_factor", C.c_float), ("soot_factor", C.c_float), ("ignition_temperature", C.c_float),
        ("burn_rate", C.c_float), ("smoke_proportion", C.c_float), ("heat_proportion", C.c_float),
        ("divergence_proportion", C.c_float),
    ]


class Source(C.Structure):
    _fields_ = [("viscosity", C.c_float), ("temperature", C.c_float), ("concentration", C.c_float),
                ("fuel", C.c_float), ("divergence", C.c_float), ("velocity_x", C.c_float),
                ("velocity_y", C.c_float), ("transfer_velocity", C.c_int32)]


SLAB_HANDLE_BYTES = 256

_lib = None


def header_symbols():
    """Every function name include/fs2d.h declares."""
    import re
    text = open(os.path.join(HERE, "..", "include", "fs2d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs2d_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libfs2d_cuda.so is not built (run `python flipsolver2d_b200/build.py`); "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    H = vp
    sig = {
        "fs2d_device_count": (i32, []),
        "fs2d_create": (i32, [C.POINTER(Params), C.POINTER(vp)]),
        "fs2d_destroy": (i32, [H]),
        "fs2d_last_error": (C.c_char_p, [H]),
        "fs2d_synchronize": (i32, [H]),
        "fs2d_stream": (vp, [H]),
        "fs2d_launch_count": (i64, [H]),
        "fs2d_grid_elements": (i64, [H, i32]),
        "fs2d_grid_element_size": (i32, [i32]),
        "fs2d_upload_grid": (i32, [H, i32, vp, C.c_size_t]),
        "fs2d_download_grid": (i32, [H, i32, vp, C.c_size_t]),
        "fs2d_grid_device_ptr": (vp, [H, i32]),
        "fs2d_clear_grid": (i32, [H, i32]),
        "fs2d_set_obstacles": (i32, [H, i32, vp]),
        "fs2d_set_sources": (i32, [H, i32, vp]),
        "fs2d_particle_count": (i64, [H]),
        "fs2d_upload_particles": (i32, [H, i64, vp, vp, vp]),
        "fs2d_download_particles": (i32, [H, vp, vp, vp]),
        "fs2d_append_particles": (i32, [H, i64, vp, vp, vp]),
        "fs2d_packed_particle_bytes": (C.c_size_t, [H, i64]),
        "fs2d_download_particles_packed": (i32, [H, vp, C.c_size_t, C.POINTER(i64)]),
        "fs2d_upload_particles_packed": (i32, [H, vp, i64]),
        "fs2d_particle_stream_bytes": (C.c_size_t, [H, i64]),
        "fs2d_particle_stream_begin": (i32, [H, vp, i64, i64]),
        "fs2d_particle_stream_positions_final": (i32, [H, vp, i64, i32]),
        "fs2d_particle_stream_velocities_final": (i32, [H, vp, i64]),
        "fs2d_particle_stream_end": (i32, [H, vp, i64, C.POINTER(i64)]),
        "fs2d_particle_stream_set_output": (i32, [H, vp, i64]),
        "fs2d_particle_stream_timing": (i32, [H, vp]),
        "fs2d_set_particle_storage_bins": (i32, [H, vp]),
        "fs2d_get_particle_storage_bins": (i32, [H, vp]),
        "fs2d_pcg_solve": (i32, [H, vp, vp, i32, f64, C.POINTER(i32)]),
        "fs2d_pcg_solve_device": (i32, [H, i32, f64]),
        "fs2d_pcg_last_iterations": (i32, [H, C.POINTER(i32)]),
        "fs2d_pcg_trace": (i32, [H, vp, i32, C.POINTER(i32)]),
        "fs2d_pcg_set_dense": (i32, [H, i32]),
        "fs2d_pcg_active_cells": (i32, [H, C.POINTER(i64)]),
        "fs2d_pcg_profile": (i32, [H, i32]),
        "fs2d_pcg_profile_read": (i32, [H, vp, vp]),
        "fs2d_spmv": (i32, [H, vp, vp]),
        "fs2d_precond_apply": (i32, [H, vp, vp]),
        "fs2d_download_matrix": (i32, [H, vp, vp, vp, vp]),
        "fs2d_set_step_dt": (i32, [H, f32]),
        "fs2d_max_particle_velocity": (i32, [H, C.POINTER(f32)]),
        "fs2d_advect": (i32, [H]),
        "fs2d_build_matrix": (i32, [H]),
        "fs2d_sort_particles": (i32, [H]),
        "fs2d_density_correction": (i32, [H, C.POINTER(i32)]),
        "fs2d_update_density_grid": (i32, [H]),
        "fs2d_density_rhs": (i32, [H]),
        "fs2d_particle_to_grid": (i32, [H]),
        "fs2d_update_sdf": (i32, [H]),
        "fs2d_update_materials": (i32, [H]),
        "fs2d_after_transfer": (i32, [H]),
        "fs2d_extrapolate_velocity": (i32, [H, i32]),
        "fs2d_extrapolate_sdf_inside": (i32, [H]),
        "fs2d_extrapolate_sdf_outside": (i32, [H]),
        "fs2d_save_velocity": (i32, [H]),
        "fs2d_apply_body_forces": (i32, [H]),
        "fs2d_pressure_rhs": (i32, [H]),
        "fs2d_apply_pressure": (i32, [H]),
        "fs2d_project": (i32, [H, C.POINTER(i32)]),
        "fs2d_velocity_from_solids": (i32, [H]),
        "fs2d_apply_viscosity": (i32, [H, C.POINTER(i32)]),
        "fs2d_particle_update": (i32, [H]),
        "fs2d_count_particles": (i32, [H]),
        "fs2d_reseed_plan": (i32, [H, C.POINTER(i64)]),
        "fs2d_reseed_apply": (i32, [H, i64, vp]),
        "fs2d_nbflip_advect_grids": (i32, [H]),
        "fs2d_set_sdf_band": (i32, [H, i32]),
        "fs2d_substep": (i32, [H, f32, vp, vp]),
        "fs2d_pcg_set_stepwise": (i32, [H, i32]),
        "fs2d_pcg_profile_solves": (i32, [H, vp, vp]),
        "fs2d_pcg_set_grid_limit": (i32, [H, i32]),
        "fs2d_pcg_set_tile_kernels": (i32, [H, i32]),
        "fs2d_pcg_set_resident": (i32, [H, i32]),
        "fs2d_pcg_last_kernel": (i32, [H, C.POINTER(C.c_int)]),
        "fs2d_state_bytes": (i32, [H, C.POINTER(C.c_size_t)]),
        "fs2d_state_save": (i32, [H, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "fs2d_state_load": (i32, [H, vp, C.c_size_t]),
        "fs2d_kernel_profile": (i32, [H, i32]),
        "fs2d_kernel_profile_read": (i32, [H, vp, vp]),
        "fs2d_slab_configure": (i32, [H, i32, i32, i32]),
        "fs2d_slab_configure_rows": (i32, [H, i32, i32, i32, vp]),
        "fs2d_slab_export": (i32, [H, vp]),
        "fs2d_slab_connect": (i32, [H, i32, vp]),
        "fs2d_slab_rows": (i32, [H, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "fs2d_slab_allgather": (i32, [H, vp, vp]),
        "fs2d_slab_gather_grid": (i32, [H, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._fs2d_signatures = sig
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Fs2dError(RuntimeError):
    pass


class Device:
    """One fs2d handle. Thin: numpy in/out, every method is one C-ABI call."""

    def __init__(self, size_i, size_j, **kw):
        self.L = lib()
        p = Params()
        p.size_i, p.size_j = int(size_i), int(size_j)
        defaults = dict(num_properties=2, particles_per_cell=8, pcg_iter_limit=200, sim_type=SIM_LIQUID,
                        parameter_handling=PARAMS_PARTICLE, viscosity_enabled=0, convergence_threads=0,
                        device=0, viscosity_property=1, temperature_property=-1, concentration_property=-1,
                        fuel_property=-1, test_property=0, dx=1.0, fluid_density=1.0, project_tolerance=1e-2,
                        gravity_x=9.8, gravity_y=0.0, pic_ratio=0.03, particle_scale=0.8,
                        ambient_temperature=273.0, temperature_decay=0.0, concentration_decay=0.0,
                        buoyancy_factor=1.0, soot_factor=1.0, ignition_temperature=250.0, burn_rate=0.05,
                        smoke_proportion=1.0, heat_proportion=1.0, divergence_proportion=0.1)
        defaults.update(kw)
        for k, v in defaults.items():
            setattr(p, k, v)
        self.params = p
        self.I, self.J = p.size_i, p.size_j
        self.N = self.I * self.J
        self.K = p.num_properties
        h = C.c_void_p()
        rc = self.L.fs2d_create(C.byref(p), C.byref(h))
        if rc != OK:
            raise Fs2dError("fs2d_create failed: %s (no CPU fallback exists)" % ERRORS.get(rc, rc))
        self.h = h

    @classmethod
    def borrow(cls, handle, size_i, size_j, num_properties=None):
        """View of a handle owned by someone else (the C++ host solver)."""
        d = cls.__new__(cls)
        d.L = lib()
        d.h = C.c_void_p(handle)
        d.I, d.J = size_i, size_j
        d.N = size_i * size_j
        d.K = num_properties
        d.params = None
        d._borrowed = True
        return d

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.L.fs2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != OK:
            msg = self.L.fs2d_last_error(self.h)
            raise Fs2dError("%s: %s (%s)" % (what, ERRORS.get(rc, rc), msg.decode() if msg else ""))

    # ---- grids
    def upload(self, name, data):
        gid, dt = GRID[name]
        a = np.ascontiguousarray(data, dt).ravel()
        self._ck(self.L.fs2d_upload_grid(self.h, gid, _p(a), a.nbytes), "upload " + name)

    def download(self, name):
        gid, dt = GRID[name]
        n = int(self.L.fs2d_grid_elements(self.h, gid))
        out = np.zeros(n, dt)
        self._ck(self.L.fs2d_download_grid(self.h, gid, _p(out), out.nbytes), "download " + name)
        return out

    def set_obstacles(self, friction):
        f = np.ascontiguousarray(friction, np.float32)
        self._ck(self.L.fs2d_set_obstacles(self.h, f.size, _p(f) if f.size else None), "set_obstacles")

    def set_sources(self, sources):
        arr = (Source * max(len(sources), 1))()
        for k, s in enumerate(sources):
            for f, _ in Source._fields_:
                setattr(arr[k], f, s[f])
        self._ck(self.L.fs2d_set_sources(self.h, len(sources), C.cast(arr, C.c_void_p) if sources else None), "set_sources")

    # ---- particles
    def particle_count(self):
        return int(self.L.fs2d_particle_count(self.h))

    def upload_particles(self, pos, vel=None, props=None, append=False):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        props = None if props is None else np.ascontiguousarray(props, np.float32)
        fn = self.L.fs2d_append_particles if append else self.L.fs2d_upload_particles
        self._ck(fn(self.h, pos.shape[0], _p(pos), _p(vel), _p(props)), "upload_particles")

    def set_storage_bins(self, bins):
        b = np.ascontiguousarray(bins, np.int32)
        self._ck(self.L.fs2d_set_particle_storage_bins(self.h, _p(b)), "set_storage_bins")

    def storage_bins(self):
        out = np.zeros(self.particle_count(), np.int32)
        self._ck(self.L.fs2d_get_particle_storage_bins(self.h, _p(out)), "storage_bins")
        return out

    def download_particles(self):
        n = self.particle_count()
        pos = np.zeros((n, 2), np.float32)
        vel = np.zeros((n, 2), np.float32)
        props = np.zeros((self.K, n), np.float32)
        self._ck(self.L.fs2d_download_particles(self.h, _p(pos), _p(vel), _p(props)), "download_particles")
        return pos, vel, props

    def download_packed(self, buf=None):
        """Packed particle state (pos | vel | props | storage byte) into `buf` (uint8 numpy array; allocated when
        None). Returns (buf, count)."""
        if buf is None:
            # records flagged dead since the last sort travel too: leave room for them
            buf = np.zeros(int(self.L.fs2d_packed_particle_bytes(self.h, self.particle_count())) * 5 // 4 + 4096, np.uint8)
        n = C.c_int64(0)
        self._ck(self.L.fs2d_download_particles_packed(self.h, _p(buf), buf.nbytes, C.byref(n)), "download_packed")
        return buf, n.value

    def stream_begin(self, buf, count, capacity):
        self._ck(self.L.fs2d_particle_stream_begin(self.h, _p(buf), int(count), int(capacity)), "stream_begin")

    def stream_positions_final(self, buf, capacity, props_final=True):
        self._ck(self.L.fs2d_particle_stream_positions_final(self.h, _p(buf), int(capacity), 1 if props_final else 0), "stream_positions_final")

    def stream_set_output(self, buf, capacity):
        self._ck(self.L.fs2d_particle_stream_set_output(self.h, _p(buf), int(capacity)), "stream_set_output")

    def stream_timing(self):
        """ms of the copies of the last streamed substep: H2D storage bytes / velocities / positions / columns, early D2H, final D2H."""
        out = np.zeros(6, np.float32)
        self._ck(self.L.fs2d_particle_stream_timing(self.h, _p(out)), "stream_timing")
        return dict(zip(("h2d_storage", "h2d_vel", "h2d_pos", "h2d_props", "d2h_early", "d2h_end"), [float(x) for x in out]))

    def stream_end(self, buf, capacity):
        n = C.c_int64(0)
        self._ck(self.L.fs2d_particle_stream_end(self.h, _p(buf), int(capacity), C.byref(n)), "stream_end")
        return n.value

    def upload_packed(self, buf, count):
        self._ck(self.L.fs2d_upload_particles_packed(self.h, _p(buf), int(count)), "upload_packed")

    # ---- PCG
    def pcg_solve(self, rhs, iter_limit, tol):
        rhs = np.ascontiguousarray(rhs, np.float64)
        x = np.zeros_like(rhs)
        it = C.c_int(0)
        self._ck(self.L.fs2d_pcg_solve(self.h, _p(rhs), _p(x), int(iter_limit), float(tol), C.byref(it)), "pcg_solve")
        return x, it.value

    def pcg_solve_device(self, iter_limit, tol):
        self._ck(self.L.fs2d_pcg_solve_device(self.h, int(iter_limit), float(tol)), "pcg_solve_device")

    def pcg_last_iterations(self):
        it = C.c_int(0)
        self._ck(self.L.fs2d_pcg_last_iterations(self.h, C.byref(it)), "pcg_last_iterations")
        return it.value

    def pcg_trace(self, max_iterations=4096):
        buf = np.zeros((max_iterations, 4), np.float64)
        n = C.c_int(0)
        self._ck(self.L.fs2d_pcg_trace(self.h, _p(buf), max_iterations, C.byref(n)), "pcg_trace")
        return buf[: n.value]

    def pcg_set_dense(self, dense=True):
        self._ck(self.L.fs2d_pcg_set_dense(self.h, 1 if dense else 0), "pcg_set_dense")

    def pcg_active_cells(self):
        n = C.c_int64(0)
        self._ck(self.L.fs2d_pcg_active_cells(self.h, C.byref(n)), "pcg_active_cells")
        return n.value

    def pcg_profile(self, enable=True):
        self._ck(self.L.fs2d_pcg_profile(self.h, 1 if enable else 0), "pcg_profile")

    def pcg_set_stepwise(self, stepwise=True):
        self._ck(self.L.fs2d_pcg_set_stepwise(self.h, 1 if stepwise else 0), "pcg_set_stepwise")

    def pcg_set_grid_limit(self, max_ctas=0):
        self._ck(self.L.fs2d_pcg_set_grid_limit(self.h, int(max_ctas)), "pcg_set_grid_limit")

    KERNEL_GROUPS = ["SORT", "P2G", "DENSITY", "SDF", "ADVECT", "G2P", "EXTRAPOLATE"]

    def kernel_profile(self, enable=True):
        self._ck(self.L.fs2d_kernel_profile(self.h, 1 if enable else 0), "kernel_profile")

    def kernel_profile_read(self):
        """{group: (device ms, calls)} accumulated since kernel_profile(True)."""
        ms = np.zeros(len(self.KERNEL_GROUPS), np.float64)
        n = np.zeros(len(self.KERNEL_GROUPS), np.int64)
        self._ck(self.L.fs2d_kernel_profile_read(self.h, _p(ms), _p(n)), "kernel_profile_read")
        return {g: (float(ms[k]), int(n[k])) for k, g in enumerate(self.KERNEL_GROUPS)}

    def set_sdf_band(self, layers):
        self._ck(self.L.fs2d_set_sdf_band(self.h, int(layers)), "set_sdf_band")

    def pcg_set_resident(self, resident=True):
        """True / 1: resident kernel, paged tiles allowed; 2: resident kernel without paging; False / 0: streaming kernel."""
        self._ck(self.L.fs2d_pcg_set_resident(self.h, int(resident)), "pcg_set_resident")

    def state_save(self):
        """The whole device state (grids + particle records) as one uint8 array (fs2d_state_save)."""
        n = C.c_size_t(0)
        self._ck(self.L.fs2d_state_bytes(self.h, C.byref(n)), "state_bytes")
        buf = np.zeros(n.value + 64, np.uint8)
        w = C.c_size_t(0)
        self._ck(self.L.fs2d_state_save(self.h, _p(buf), buf.nbytes, C.byref(w)), "state_save")
        return buf[: w.value]

    def state_load(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        self._ck(self.L.fs2d_state_load(self.h, _p(blob), blob.nbytes), "state_load")

    def pcg_last_kernel(self):
        """0 = streaming whole-solve kernel (or stepwise), 1 = resident, 2 = resident + paged tiles."""
        k = C.c_int(0)
        self._ck(self.L.fs2d_pcg_last_kernel(self.h, C.byref(k)), "pcg_last_kernel")
        return k.value

    def pcg_set_tile_kernels(self, tile=True):
        self._ck(self.L.fs2d_pcg_set_tile_kernels(self.h, 1 if tile else 0), "pcg_set_tile_kernels")

    def pcg_profile_solves(self):
        ms = np.zeros(1, np.float64)
        n = np.zeros(1, np.int64)
        self._ck(self.L.fs2d_pcg_profile_solves(self.h, _p(ms), _p(n)), "pcg_profile_solves")
        return float(ms[0]), int(n[0])

    def pcg_profile_read(self):
        ms = np.zeros(2, np.float64)
        n = np.zeros(2, np.int64)
        self._ck(self.L.fs2d_pcg_profile_read(self.h, _p(ms), _p(n)), "pcg_profile_read")
        return ms, n

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        self._ck(self.L.fs2d_spmv(self.h, _p(x), _p(y)), "spmv")
        return y

    def precond(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        self._ck(self.L.fs2d_precond_apply(self.h, _p(x), _p(y)), "precond")
        return y

    def matrix(self):
        n = self.N
        is_unit = np.zeros(n, np.uint8)
        mask = np.zeros(n, np.uint8)
        count = np.zeros(n, np.uint8)
        pre = np.zeros((4, n), np.uint8)
        self._ck(self.L.fs2d_download_matrix(self.h, _p(is_unit), _p(mask), _p(count), _p(pre)), "matrix")
        return dict(is_unit=is_unit, mask=mask, count=count, precond_counts=pre)

    # ---- stages
    def set_step_dt(self, dt):
        self._ck(self.L.fs2d_set_step_dt(self.h, float(dt)), "set_step_dt")

    def max_particle_velocity(self):
        v = C.c_float(0)
        self._ck(self.L.fs2d_max_particle_velocity(self.h, C.byref(v)), "max_particle_velocity")
        return v.value

    def stage(self, name, *args):
        fn = getattr(self.L, "fs2d_" + name)
        self._ck(fn(self.h, *args), name)

    def stage_iters(self, name):
        it = C.c_int(0)
        self._ck(getattr(self.L, "fs2d_" + name)(self.h, C.byref(it)), name)
        return it.value

    def reseed(self, uniform_xy_fn):
        """uniform_xy_fn(count) -> float32 array of 2*count uniforms drawn in stream order."""
        n = C.c_int64(0)
        self._ck(self.L.fs2d_reseed_plan(self.h, C.byref(n)), "reseed_plan")
        u = np.ascontiguousarray(uniform_xy_fn(n.value), np.float32) if n.value else np.zeros(0, np.float32)
        self._ck(self.L.fs2d_reseed_apply(self.h, n.value, _p(u) if n.value else None), "reseed_apply")
        return n.value

    def substep(self, dt):
        ms = np.zeros(12, np.float32)
        iters = np.zeros(3, np.int32)
        self._ck(self.L.fs2d_substep(self.h, float(dt), _p(ms), _p(iters)), "substep")
        return ms, iters

    # ---- row slabs over several GPUs (or several ranks on one GPU, for tests)
    def slab_configure(self, rank, world, device_share=1, row_bounds=None):
        if row_bounds is None:
            self._ck(self.L.fs2d_slab_configure(self.h, int(rank), int(world), int(device_share)), "slab_configure")
        else:
            b = np.ascontiguousarray(row_bounds, np.int32)
            assert b.size == world + 1
            self._ck(self.L.fs2d_slab_configure_rows(self.h, int(rank), int(world), int(device_share), _p(b)), "slab_configure_rows")
        self.rank, self.world = int(rank), int(world)

    def slab_export(self):
        buf = C.create_string_buffer(SLAB_HANDLE_BYTES)
        self._ck(self.L.fs2d_slab_export(self.h, C.cast(buf, C.c_void_p)), "slab_export")
        return bytes(buf.raw)

    def slab_connect(self, peer_rank, blob):
        buf = C.create_string_buffer(bytes(blob), SLAB_HANDLE_BYTES)
        self._ck(self.L.fs2d_slab_connect(self.h, int(peer_rank), C.cast(buf, C.c_void_p)), "slab_connect")

    def slab_rows(self):
        a, b, h = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.fs2d_slab_rows(self.h, C.byref(a), C.byref(b), C.byref(h)), "slab_rows")
        return a.value, b.value, h.value

    def slab_allgather(self, values):
        v = np.zeros(4, np.int64)
        v[: len(values)] = values
        out = np.zeros(4 * getattr(self, "world", 1), np.int64)
        self._ck(self.L.fs2d_slab_allgather(self.h, _p(v), _p(out)), "slab_allgather")
        return out.reshape(-1, 4)

    def slab_gather(self, name):
        """Collective: afterwards every rank holds all rows of the grid (and the lazily extrapolated FLUID_SDF)."""
        self._ck(self.L.fs2d_slab_gather_grid(self.h, GRID[name][0]), "slab_gather " + name)

    def synchronize(self):
        self._ck(self.L.fs2d_synchronize(self.h), "synchronize")

    def launch_count(self):
        return int(self.L.fs2d_launch_count(self.h))


def connect_slabs(devices):
    """Connect handles that live in THIS process (one per GPU, or several sharing a GPU in tests):
    configure must have been called on each; every handle maps every other one."""
    blobs = [d.slab_export() for d in devices]
    for d in devices:
        for r, blob in enumerate(blobs):
            if r != d.rank:
                d.slab_connect(r, blob)


def run_ranks(fns):
    """Run one callable per rank concurrently (ctypes releases the GIL inside the C calls; the ranks
    spin on each other's device flags, so they must not be serialised). Returns the results in rank order."""
    import threading
    out, err = [None] * len(fns), [None] * len(fns)

    def work(k):
        try:
            out[k] = fns[k]()
        except BaseException as e:  # noqa: BLE001
            err[k] = e

    ts = [threading.Thread(target=work, args=(k,)) for k in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    failed = [(k, e) for k, e in enumerate(err) if e is not None]
    if failed:
        # a rank that gives up makes the others time out: show every rank's error, not only the first in rank order
        k0, e0 = failed[0]
        if len(failed) > 1:
            raise type(e0)("; ".join("rank %d: %s" % (k, e) for k, e in failed)) from e0
        raise e0
    return out
