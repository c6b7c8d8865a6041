"""ctypes binding of include/fs2d_host.h (libfs2d_host.so): the C++ host mirror of the reference's
JsonSceneReader / FlipSolver API. bench.py and the tests drive scenes through this -- the same objects a
C++ application links against. No CPU fallback: stepping needs libfs2d_cuda.so and a CUDA device."""
import ctypes as C
import os

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfs2d_host.so")

_lib = None


def header_symbols():
    import re
    text = open(os.path.join(HERE, "..", "include", "fs2d_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs2dh_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libfs2d_host.so is not built (run `python flipsolver2d_b200/build_host.py`)")
    capi.lib()  # libfs2d_cuda.so first: the host library links against it
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    sig = {
        "fs2dh_set_quiet": (None, [i32]),
        "fs2dh_set_device": (None, [i32]),
        "fs2dh_set_convergence_threads": (None, [i32]),
        "fs2dh_set_slab": (None, [i32, i32, i32]),
        "fs2dh_slab_export": (i32, [vp, vp]),
        "fs2dh_slab_connect": (i32, [vp, i32, vp]),
        "fs2dh_global_particle_count": (i64, [vp]),
        "fs2dh_slab_bounds": (i32, [vp, i32, vp]),
        "fs2dh_write_stats_xlsx": (i32, [C.c_char_p, i32, vp, vp, vp]),
        "fs2dh_load_scene": (vp, [C.c_char_p]),
        "fs2dh_destroy": (None, [vp]),
        "fs2dh_last_error": (C.c_char_p, [vp]),
        "fs2dh_prepare_host": (i32, [vp]),
        "fs2dh_seed_count": (i64, [vp]),
        "fs2dh_seed_particles": (i32, [vp, vp, vp, vp]),
        "fs2dh_host_grid": (i32, [vp, i32, vp]),
        "fs2dh_prepare": (i32, [vp]),
        "fs2dh_step_frame": (i32, [vp]),
        "fs2dh_step_substep": (i32, [vp, C.POINTER(i32)]),
        "fs2dh_step_substep_streamed": (i32, [vp, vp, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(i32)]),
        "fs2dh_save_state": (i32, [vp, C.c_char_p]),
        "fs2dh_load_state": (i32, [vp, C.c_char_p]),
        "fs2dh_get_stats": (i32, [vp, vp, vp]),
        "fs2dh_size_i": (i32, [vp]),
        "fs2dh_size_j": (i32, [vp]),
        "fs2dh_sim_type": (i32, [vp]),
        "fs2dh_frame_number": (i32, [vp]),
        "fs2dh_particle_count": (i64, [vp]),
        "fs2dh_kernel_launches": (i64, [vp]),
        "fs2dh_device": (vp, [vp]),
        "fs2dh_material": (i32, [vp, vp]),
        "fs2dh_bin_sizes": (i64, [vp, vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def write_stats_xlsx(path, scenes):
    """scenes: {name: array of shape (frames, 18)} -> Stats.xlsx as AutoBench's BenchRunTable::save lays it out."""
    L = lib()
    names = list(scenes.keys())
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    counts = np.array([len(scenes[n]) for n in names], np.int32)
    rows = np.ascontiguousarray(np.concatenate([np.asarray(scenes[n], np.float64).reshape(-1, 18) for n in names])
                                if names else np.zeros((0, 18)), np.float64)
    rc = L.fs2dh_write_stats_xlsx(str(path).encode(), len(names), C.cast(arr, C.c_void_p), _p(counts), _p(rows))
    if rc != 0:
        raise capi.Fs2dError("write_stats_xlsx failed (%d)" % rc)


def connect_slabs(solvers):
    """Connect Solver objects living in THIS process (tests: several ranks on one GPU)."""
    blobs = [s.slab_export() for s in solvers]
    for s in solvers:
        for r, blob in enumerate(blobs):
            if r != s.rank:
                s.slab_connect(r, blob)


STAGES = ["ADVECTION", "DECOMPOSITION", "DENSITY", "PARTICLE_REBIN", "PARTICLE_TO_GRID", "GRID_UPDATE", "AFTER_TRANSFER",
          "PRESSURE", "VISCOSITY", "REPRESSURE", "PARTICLE_UPDATE", "PARTICLE_RESEED"]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Solver:
    """JsonSceneReader::loadJson + FlipSolver::stepFrame through the C shim."""

    def __init__(self, json_path, quiet=True, device=0, convergence_threads=0, slab=None):
        """slab = (rank, world[, device_share]): this solver owns one row slab of the scene (one process per GPU;
        device_share > 1 only for tests that run several ranks on one GPU). The device handle is created here so
        that the process-wide slab setting cannot leak into a solver made later."""
        self.L = lib()
        self.L.fs2dh_set_quiet(1 if quiet else 0)
        self.L.fs2dh_set_device(int(device))
        self.L.fs2dh_set_convergence_threads(int(convergence_threads))
        self.rank, self.world = 0, 1
        if slab is not None:
            self.rank, self.world = int(slab[0]), int(slab[1])
            self.L.fs2dh_set_slab(self.rank, self.world, int(slab[2]) if len(slab) > 2 else 1)
        else:
            self.L.fs2dh_set_slab(0, 1, 1)
        self.h = self.L.fs2dh_load_scene(str(json_path).encode())
        if not self.h:
            raise RuntimeError("JsonSceneReader::loadJson failed for %s" % json_path)
        self.I = self.L.fs2dh_size_i(self.h)
        self.J = self.L.fs2dh_size_j(self.h)
        self.N = self.I * self.J
        if slab is not None:
            if not self.L.fs2dh_device(self.h):
                raise capi.Fs2dError("no device: %s" % self.L.fs2dh_last_error(self.h).decode())
            self.L.fs2dh_set_slab(0, 1, 1)

    def slab_export(self):
        buf = C.create_string_buffer(capi.SLAB_HANDLE_BYTES)
        self._ck(self.L.fs2dh_slab_export(self.h, C.cast(buf, C.c_void_p)), "slab_export")
        return bytes(buf.raw)

    def slab_connect(self, peer_rank, blob):
        buf = C.create_string_buffer(bytes(blob), capi.SLAB_HANDLE_BYTES)
        self._ck(self.L.fs2dh_slab_connect(self.h, int(peer_rank), C.cast(buf, C.c_void_p)), "slab_connect")

    def slab_bounds(self, world):
        out = np.zeros(world + 1, np.int32)
        self._ck(self.L.fs2dh_slab_bounds(self.h, int(world), _p(out)), "slab_bounds")
        return out

    def global_particle_count(self):
        n = int(self.L.fs2dh_global_particle_count(self.h))
        if n < 0:
            raise capi.Fs2dError("global_particle_count: %s" % self.L.fs2dh_last_error(self.h).decode())
        return n

    def close(self):
        if getattr(self, "h", None):
            self.L.fs2dh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise capi.Fs2dError("%s: %s" % (what, self.L.fs2dh_last_error(self.h).decode()))

    def prepare_host(self):
        self._ck(self.L.fs2dh_prepare_host(self.h), "prepare_host")

    def prepare(self):
        self._ck(self.L.fs2dh_prepare(self.h), "prepare")

    def step_frame(self):
        self._ck(self.L.fs2dh_step_frame(self.h), "step_frame")

    def step_substep(self):
        done = C.c_int(0)
        self._ck(self.L.fs2dh_step_substep(self.h, C.byref(done)), "step_substep")
        return bool(done.value)

    def step_substep_streamed(self, host_ptr, capacity_records, count_in):
        """One substep with the particle state in a (pinned) host buffer, sectioned layout (include/fs2d.h).
        Returns (frame_finished, records now in the buffer)."""
        done = C.c_int(0)
        n = C.c_int64(0)
        self._ck(self.L.fs2dh_step_substep_streamed(self.h, C.c_void_p(host_ptr), int(capacity_records), int(count_in), C.byref(n),
                                                    C.byref(done)), "step_substep_streamed")
        return bool(done.value), n.value

    def save_state(self, path):
        """FlipSolver::saveState: device grids + particle records + frame / substep counters + mt19937 stream."""
        self._ck(self.L.fs2dh_save_state(self.h, str(path).encode()), "save_state")

    def load_state(self, path):
        self._ck(self.L.fs2dh_load_state(self.h, str(path).encode()), "load_state")

    def stats(self):
        t = np.zeros(12, np.float32)
        m = np.zeros(5, np.float32)
        self._ck(self.L.fs2dh_get_stats(self.h, _p(t), _p(m)), "get_stats")
        return dict(timings=t, frame_ms=float(m[0]), substeps=int(m[1]), pressure_iters=int(m[2]),
                    density_iters=int(m[3]), viscosity_iters=int(m[4]))

    def particle_count(self):
        return int(self.L.fs2dh_particle_count(self.h))

    def kernel_launches(self):
        return int(self.L.fs2dh_kernel_launches(self.h))

    def frame_number(self):
        return self.L.fs2dh_frame_number(self.h)

    def seed_particles(self, num_props):
        n = int(self.L.fs2dh_seed_count(self.h))
        pos = np.zeros((n, 2), np.float32)
        vel = np.zeros((n, 2), np.float32)
        props = np.zeros((num_props, n), np.float32)
        self._ck(self.L.fs2dh_seed_particles(self.h, _p(pos), _p(vel), _p(props)), "seed_particles")
        return pos, vel, props

    def host_grid(self, name):
        gid, dt = capi.GRID[name]
        out = np.zeros(self.N, dt)
        self._ck(self.L.fs2dh_host_grid(self.h, gid, _p(out)), "host_grid " + name)
        return out

    def material(self):
        out = np.zeros(self.N, np.int8)
        self._ck(self.L.fs2dh_material(self.h, _p(out)), "material")
        return out

    def bin_sizes(self):
        n = ((self.I + 2) // 3) * ((self.J + 2) // 3)
        out = np.zeros(n, np.int32)
        got = self.L.fs2dh_bin_sizes(self.h, _p(out), n)
        assert got == n
        return out

    def device(self, num_properties=None):
        """A capi.Device view of the solver's fs2d handle (not owned)."""
        h = self.L.fs2dh_device(self.h)
        if not h:
            raise capi.Fs2dError("no device: %s" % self.L.fs2dh_last_error(self.h).decode())
        return capi.Device.borrow(h, self.I, self.J, num_properties)
