"""Kernel group 4 parity: CUDA PCG (through the C ABI) vs the reference's own LinearSolver,
IndexedPressureParameters and IndexedIPPCoefficients compiled into oracle/_ref (strict build)."""
import numpy as np
import pytest

from conftest import ORACLE_THREADS
from flipsolver2d_b200 import capi, scenes

import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _ref_system(ref_mod, scene_dir, res, sim="flip", frames=0, dt=1.0 / 90.0):
    """Reference solver advanced to the point where step() builds the matrix."""
    path = scenes.write_scene(scenes.dam_break(res, sim), str(scene_dir / ("db%d_%s.json" % (res, sim))))
    s = ref_mod.RefSolver(path, strict=True)
    for _ in range(frames):
        s.step_frame()
    if frames == 0:
        s.stage("FIRST_FRAME_INIT")
    s.set_step_dt(dt)
    s.stage("BUILD_MATRIX")
    return s


def _device_for(s, conv_threads=0, iter_limit=200):
    p = s.params()
    d = capi.Device(s.I, s.J, dx=p["dx"], fluid_density=p["fluidDensity"], pcg_iter_limit=iter_limit,
                    convergence_threads=conv_threads, project_tolerance=p["projectTolerance"])
    d.upload("MATERIAL", s.grid("MATERIAL"))
    d.set_step_dt(p["stepDt"])
    d.stage("build_matrix")
    return d


@pytest.mark.parametrize("res,frames", [(64, 0), (128, 2), (100, 1)])
def test_matrix_rows_bit_exact(ref_mod, scene_dir, res, frames):
    s = _ref_system(ref_mod, scene_dir, res, frames=frames)
    d = _device_for(s)
    rm, dm = s.matrix(), d.matrix()
    assert np.array_equal(rm["is_unit"], dm["is_unit"])
    assert np.array_equal(rm["mask"], dm["mask"])
    assert np.array_equal(rm["count"], dm["count"])
    unit = rm["is_unit"].astype(bool)
    scale = rm["scale"]
    with np.errstate(divide="ignore"):
        coef = 1.0 / (dm["precond_counts"].astype(np.float64) * scale)
    for k in range(4):
        assert np.array_equal(coef[k][unit], rm["coef"][k][unit])


@pytest.mark.parametrize("res,frames", [(64, 0), (128, 2), (100, 1)])
def test_operators_bit_exact(ref_mod, scene_dir, res, frames):
    s = _ref_system(ref_mod, scene_dir, res, frames=frames)
    d = _device_for(s)
    rng = np.random.default_rng(res)
    v = rng.standard_normal(s.N)
    assert np.array_equal(d.spmv(v), s.spmv(v))
    assert np.array_equal(d.precond(v), s.precond(v))


@pytest.mark.parametrize("res,frames,iters", [(64, 0, 10), (128, 2, 25), (256, 1, 40)])
def test_fixed_iteration_solve(ref_mod, scene_dir, res, frames, iters):
    """tol = 0 forces exactly `iters` iterations on both sides; compare the iterate."""
    s = _ref_system(ref_mod, scene_dir, res, frames=frames)
    d = _device_for(s)
    rng = np.random.default_rng(7)
    unit = s.matrix()["is_unit"].astype(bool)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    xr, itr = s.pcg(rhs, iters, 0.0)
    xd, itd = d.pcg_solve(rhs, iters, 0.0)
    assert itr == itd == iters
    rel = np.linalg.norm(xd - xr) / np.linalg.norm(xr)
    # tolerance: 1e-9 relative L2 on the iterate (order of the dot-product sums differs)
    assert rel < 1e-9, rel


@pytest.mark.parametrize("res,frames,iters", [(256, 1, 40), (384, 0, 30)])
def test_active_tile_walk_equals_dense_walk(ref_mod, scene_dir, res, frames, iters):
    """The iteration kernels skip tiles without matrix rows and with a zero right-hand side; the iterate
    must be the one the dense walk produces (only the grouping of the dot-product partials differs) and
    fewer cells must be walked."""
    s = _ref_system(ref_mod, scene_dir, res, frames=frames)
    d = _device_for(s)
    rng = np.random.default_rng(11)
    unit = s.matrix()["is_unit"].astype(bool)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    rhs[5 * s.J + 7] = 0.25  # a non-zero rhs on an identity row far from the fluid keeps its tile active
    d.pcg_set_dense(False)
    xa, ita = d.pcg_solve(rhs, iters, 0.0)
    active = d.pcg_active_cells()
    d.pcg_set_dense(True)
    xd, itd = d.pcg_solve(rhs, iters, 0.0)
    assert ita == itd == iters
    assert 0 < active < s.N
    assert np.linalg.norm(xa - xd) <= 1e-10 * np.linalg.norm(xd)
    xr, _ = s.pcg(rhs, iters, 0.0)
    assert np.linalg.norm(xa - xr) / np.linalg.norm(xr) < 1e-9
    # untouched tiles hold exact zeros, the lone identity row its rhs (x = rhs after one step of identity)
    assert not xa[~unit & (rhs == 0)].any()
    d.close()


def test_zero_rhs_and_iteration_count_compat(ref_mod, scene_dir):
    s = _ref_system(ref_mod, scene_dir, 256, frames=1, dt=1.0 / 30.0)  # scale 1.75: SPD regime
    assert s.threads == ORACLE_THREADS
    d = _device_for(s, conv_threads=ORACLE_THREADS, iter_limit=400)
    x, it = d.pcg_solve(np.zeros(s.N), 50, 1e-6)
    assert it == 0 and not x.any()
    rhs = s.pressure_rhs()
    xr, itr = s.pcg(rhs, 400, 1e-2)
    xd, itd = d.pcg_solve(rhs, 400, 1e-2)
    assert itr < 400, "expected a converging case"
    assert itd == itr
    rel = np.linalg.norm(xd - xr) / np.linalg.norm(xr)
    assert rel < 1e-7, rel


def test_golden_matviz_system_through_cuda(ref_mod, scene_dir):
    """The reference's own dumped pressure system (tests/golden, see test_oracle_golden.py) through the
    CUDA operator and solver: A*vout == vin to the printed digits, PCG(vin) == vout."""
    import os
    from test_oracle_golden import GOLDEN, golden_material
    g = np.load(GOLDEN)
    I, J = (int(v) for v in g["size"])
    d = capi.Device(I, J, dx=50.0 / 64, fluid_density=0.5, pcg_iter_limit=2000)
    d.upload("MATERIAL", golden_material(g))
    d.set_step_dt(1.0 / 30.0)
    d.stage("build_matrix")
    m = d.matrix()
    idx = g["index"]
    assert np.array_equal(np.nonzero(m["is_unit"])[0], idx)
    assert np.abs(0.109227 * m["count"][idx] - g["diag"]).max() < 1e-5
    assert np.abs(d.spmv(g["vout"]) - g["vin"]).max() < 1e-4
    x, iters = d.pcg_solve(g["vin"], 2000, 1e-6)
    assert iters < 2000
    assert np.abs(x - g["vout"]).max() < 1e-3 * np.abs(g["vout"]).max()
    d.close()


def test_iteration_count_compat_eight_threads(tmp_path):
    """The reference's convergence value depends on its ThreadPool size (vmath.cpp:100-136); the suite
    pins the oracle to one thread, so the T = 8 case runs in its own process."""
    import os
    import subprocess
    import sys
    code = r"""
import os, sys
os.environ["FS2D_ORACLE_THREADS"] = "8"
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from oracle import ref
from flipsolver2d_b200 import capi, scenes
path = scenes.write_scene(scenes.dam_break(256, "flip"), %r)
s = ref.RefSolver(path, strict=True, threads=8)
s.step_frame()
s.set_step_dt(1.0 / 30.0)
s.stage("BUILD_MATRIX")
assert s.threads == 8
p = s.params()
d = capi.Device(s.I, s.J, dx=p["dx"], fluid_density=p["fluidDensity"], pcg_iter_limit=400, convergence_threads=8)
d.upload("MATERIAL", s.grid("MATERIAL")); d.set_step_dt(p["stepDt"]); d.stage("build_matrix")
rhs = s.pressure_rhs()
xr, itr = s.pcg(rhs, 400, 1e-2)
xd, itd = d.pcg_solve(rhs, 400, 1e-2)
d1 = capi.Device(s.I, s.J, dx=p["dx"], fluid_density=p["fluidDensity"], pcg_iter_limit=400, convergence_threads=1)
d1.upload("MATERIAL", s.grid("MATERIAL")); d1.set_step_dt(p["stepDt"]); d1.stage("build_matrix")
_, it1 = d1.pcg_solve(rhs, 400, 1e-2)
print("ITERS", itr, itd, it1, float(np.linalg.norm(xd - xr) / np.linalg.norm(xr)))
assert itr < 400 and itd == itr
""" % (ROOT, ROOT, str(tmp_path / "t8.json"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ITERS" in r.stdout


@pytest.mark.parametrize("dense", [False, True])
def test_whole_solve_kernel_equals_stepwise_kernels(ref_mod, scene_dir, dense):
    """The default solve is ONE cooperative launch (pcgSolveKernel); the two-kernels-per-iteration path does the same
    arithmetic with the same tile -> CTA assignment and reduction order, so iterates agree to the last bit."""
    s = _ref_system(ref_mod, scene_dir, 192, dt=1.0 / 2000.0)  # small matrix scale: the solve converges within the cap
    d = _device_for(s, iter_limit=400)
    d.pcg_set_dense(dense)
    d.pcg_set_resident(False)  # the streaming whole-solve kernel; pcgResidentKernel groups the partials differently
    unit = d.matrix()["is_unit"].astype(bool)
    rng = np.random.default_rng(11)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    for limit, tol in ((37, 0.0), (300, 1e-7), (5, 1e30), (0, 0.0)):
        d.pcg_set_stepwise(True)
        x_step, n_step = d.pcg_solve(rhs, limit, tol)
        t_step = d.pcg_trace().copy()
        d.pcg_set_stepwise(False)
        x_whole, n_whole = d.pcg_solve(rhs, limit, tol)
        t_whole = d.pcg_trace().copy()
        assert n_whole == n_step
        assert np.array_equal(t_whole, t_step)
        assert np.array_equal(x_whole, x_step)
    x0, n0 = d.pcg_solve(np.zeros(s.N), 50, 0.0)  # VOps::isZero: no iterations
    assert n0 == 0 and not x0.any()
    d.close()


def test_linear_index_wrap_columns_without_walls(ref_mod, scene_dir):
    """A tank without side walls: FLUID cells in columns 0 and J-1 have their j-1 / j+1 "neighbour" bit set (the material
    grid extends its border, materialgrid.cpp:5-8) and the reference's operators then read the LINEAR neighbours idx-1 /
    idx+1, i.e. the last / first element of the adjacent row (pressuredata.h:135-145). The whole-solve kernel loads its
    tiles with tensor copies, which zero-fill outside the matrix, and patches exactly those wrap columns by hand; the
    stepwise kernels load row segments by linear index. Both must reproduce the reference's iterates."""
    scene = scenes.dam_break(96, "flip")
    scene["solver"]["objects"] = [o for o in scene["solver"]["objects"] if o["type"] != "solid"]
    scene["solver"]["objects"][-1]["verts"] = [[20, 0], [20, 50], [45, 50], [45, 0]]  # fluid across the full width
    path = scenes.write_scene(scene, str(scene_dir / "nowalls.json"))
    s = ref_mod.RefSolver(path, strict=True)
    s.stage("FIRST_FRAME_INIT")
    s.set_step_dt(1.0 / 2000.0)
    s.stage("BUILD_MATRIX")
    d = _device_for(s, iter_limit=60)
    rm = s.matrix()
    unit = rm["is_unit"].astype(bool).reshape(s.I, s.J)
    assert unit[:, 0].any() and unit[:, -1].any()            # rows exist in the first and the last column ...
    mask = rm["mask"].reshape(s.I, s.J)
    assert (mask[unit[:, 0], 0] & 4).any() and (mask[unit[:, -1], -1] & 8).any()  # ... with the wrap neighbours switched on
    rng = np.random.default_rng(3)
    rhs = np.where(unit.ravel(), rng.standard_normal(s.N), 0.0)
    v = rng.standard_normal(s.N)
    assert np.array_equal(d.spmv(v), s.spmv(v))
    for dense in (True, False):
        d.pcg_set_dense(dense)
        d.pcg_set_resident(False)
        xr, nr = s.pcg(rhs, 25, 0.0)
        d.pcg_set_stepwise(True)
        x_step, n_step = d.pcg_solve(rhs, 25, 0.0)
        d.pcg_set_stepwise(False)
        x_whole, n_whole = d.pcg_solve(rhs, 25, 0.0)
        assert n_step == n_whole == nr == 25
        assert np.array_equal(x_whole, x_step)
        assert np.linalg.norm(x_whole - xr) / np.linalg.norm(xr) < 1e-9
        if not dense:
            # the resident kernel reads the wrap neighbours as ring cells by linear index
            d.pcg_set_resident(True)
            x_res, n_res = d.pcg_solve(rhs, 25, 0.0)
            assert n_res == 25
            assert np.linalg.norm(x_res - xr) / np.linalg.norm(xr) < 1e-9
    d.close()
    s.close()
