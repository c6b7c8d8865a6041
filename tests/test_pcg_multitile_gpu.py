"""PCG parity where the benchmark lives: every CTA of the persistent kernels walks SEVERAL tiles.

At 4096^2 a CTA of pcgSolveKernel walks ~28 tiles through the double-buffered TMA pipeline (prefetch of tile k+1 into
the other shared-memory stage, mbarrier parity bookkeeping, stage reuse after fence.proxy.async, the early half of the
first tile issued before the grid barrier). The small systems of test_pcg_gpu.py give every CTA at most one tile, so
none of that runs there. Here it does, two ways:
  * fs2d_pcg_set_grid_limit shrinks the persistent grid, so 256^2 (32 tiles) walks 4 .. 16 tiles per CTA, odd and even
    counts, ragged last round;
  * 1024^2 flip (512 tiles) and 2048^2 smoke rows (2048 tiles, all non-solid cells are DOFs) run on the full grid of
    2 x 148 CTAs and on a shrunk one.
Every case is compared with the reference's own LinearSolver (oracle/_ref, strict build) at 1e-9 relative L2 on the
iterate, and the three device evaluations of the same arithmetic -- whole-solve kernel, two kernels per iteration,
plain tile kernels -- with each other (linearsolver.cpp:25-73, pressuredata.h:132-238, PressureIPPCoeficients.h:23-132).
"""
import numpy as np
import pytest

import helpers as H
from flipsolver2d_b200 import capi, scenes

pytestmark = pytest.mark.gpu

TOL = 1e-9  # relative L2 of the iterate vs the reference: only the grouping of the dot-product partials differs


def _system(ref_mod, scene_dir, scene, name, dt, frames=0):
    path = scenes.write_scene(scene, str(scene_dir / (name + ".json")))
    s = ref_mod.RefSolver(path, strict=True)
    for _ in range(frames):
        s.step_frame()
    if frames == 0:
        s.stage("FIRST_FRAME_INIT")
    s.set_step_dt(dt)
    s.stage("BUILD_MATRIX")
    return s


def _device(s, scene, iter_limit=200):
    d = H.make_device(s, scene, pcg_iter_limit=iter_limit)
    d.upload("MATERIAL", s.grid("MATERIAL"))
    d.set_step_dt(s.params()["stepDt"])
    d.stage("build_matrix")
    return d


def _rhs(s, seed, lone=True):
    rng = np.random.default_rng(seed)
    unit = s.matrix()["is_unit"].astype(bool)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    if lone:
        rhs[5 * s.J + 7] = 0.25  # non-zero rhs on an identity row far from the fluid: one more active tile
    return rhs, unit


def _modes(d, rhs, iters, tol=0.0):
    """whole-solve kernel, stepwise kernels, plain tile kernels: iterate + iteration count + trace of each."""
    out = {}
    d.pcg_set_tile_kernels(False)
    d.pcg_set_stepwise(False)
    d.pcg_set_resident(True)   # active walk with few enough tiles: pcgResidentKernel (vectors stay in shared memory)
    out["resident"] = d.pcg_solve(rhs, iters, tol) + (d.pcg_trace().copy(),)
    d.pcg_set_resident(False)  # the streaming whole-solve kernel
    out["whole"] = d.pcg_solve(rhs, iters, tol) + (d.pcg_trace().copy(),)
    d.pcg_set_stepwise(True)
    out["stepwise"] = d.pcg_solve(rhs, iters, tol) + (d.pcg_trace().copy(),)
    d.pcg_set_stepwise(False)
    d.pcg_set_tile_kernels(True)
    out["tile"] = d.pcg_solve(rhs, iters, tol) + (d.pcg_trace().copy(),)
    d.pcg_set_tile_kernels(False)
    d.pcg_set_resident(True)
    return out


@pytest.fixture(scope="module")
def sys256(ref_mod, scene_dir):
    scene = scenes.dam_break(256, "flip")
    s = _system(ref_mod, scene_dir, scene, "mt256", 1.0 / 90.0, frames=1)
    d = _device(s, scene)
    yield s, d
    d.close()
    s.close()


@pytest.mark.parametrize("dense", [True, False], ids=["dense", "active"])
@pytest.mark.parametrize("limit", [2, 3, 5, 7])
def test_shrunk_grid_walks_many_tiles_per_cta(sys256, limit, dense):
    """256^2 = 32 tiles of 16x128. limit 2 -> 16 tiles per CTA (even), 3 -> 11/11/10, 5 -> 7/7/6/6/6, 7 -> 5/5/5/5/4/4/4;
    the active walk (tiles with matrix rows or a non-zero rhs) leaves fewer, ragged lists."""
    s, d = sys256
    rhs, unit = _rhs(s, 100 + limit)
    iters = 30
    xr, itr = s.pcg(rhs, iters, 0.0)
    d.pcg_set_dense(dense)
    d.pcg_set_grid_limit(limit)
    try:
        m = _modes(d, rhs, iters)
        if not dense:
            active_tiles = d.pcg_active_cells() // (16 * 128)
            assert limit < active_tiles < 32, active_tiles  # really several tiles per CTA, really fewer than all
    finally:
        d.pcg_set_grid_limit(0)
        d.pcg_set_dense(False)
    for name, (x, it, tr) in m.items():
        assert it == itr == iters, name
        assert H.rel_l2(x, xr) < TOL, (name, H.rel_l2(x, xr))
    # same tile -> CTA assignment, same reduction order: the two pipelined evaluations agree to the last bit
    assert np.array_equal(m["whole"][0], m["stepwise"][0])
    assert np.array_equal(m["whole"][2], m["stepwise"][2])
    # the plain tile kernels group the partials per tile: same numbers up to that regrouping
    assert H.rel_l2(m["tile"][0], m["whole"][0]) < 1e-11
    assert H.rel_l2(m["resident"][0], m["whole"][0]) < 1e-11
    if not dense:
        assert not m["whole"][0][~unit & (rhs == 0)].any()  # skipped tiles hold exact zeros
        assert not m["resident"][0][~unit & (rhs == 0)].any()


def test_shrunk_grid_converging_solve_and_zero_rhs(sys256):
    """Convergence exit in the middle of a multi-tile walk (early-issued half tile in flight), then a zero rhs."""
    s, d = sys256
    scene = scenes.dam_break(256, "flip")
    s2 = None
    d.pcg_set_grid_limit(3)
    try:
        rhs, _ = _rhs(s, 5)
        # tolerance picked from the trace so that the loop exits after ~12 iterations on every path
        d.pcg_solve(rhs, 30, 0.0)
        tr = d.pcg_trace()
        tol = float(tr[12, 3]) * 1.0000001
        first = int(np.argmax(tr[:, 3] <= tol))
        m = _modes(d, rhs, 30, tol)
        for name, (x, it, t) in m.items():
            assert it == first, (name, it, first)
        assert np.array_equal(m["whole"][0], m["stepwise"][0])
        x0, n0 = d.pcg_solve(np.zeros(s.N), 30, 0.0)
        assert n0 == 0 and not x0.any()
        # and the kernel is reusable after the early exit
        x1, n1 = d.pcg_solve(rhs, 30, 0.0)
        xr, _ = s.pcg(rhs, 30, 0.0)
        assert n1 == 30 and H.rel_l2(x1, xr) < TOL
    finally:
        d.pcg_set_grid_limit(0)
    del scene, s2


@pytest.mark.parametrize("per_cta", [5, 7, 11])
def test_paged_resident_tiles(sys256, per_cta):
    """More active tiles than the CTAs hold resident (4 each), up to 12 per CTA: pcgResidentKernel<PAGED> keeps 3 tiles in
    shared memory and pages the private boxes of the others through two scratch boxes with bulk copies (the 4096^2 dam
    break on one GPU: 904 tiles for 148 SMs). per_cta 5 -> 3 resident + 2 paged (one box prefetched across every phase
    boundary), 7 -> + 4 paged (double-buffered inside a phase, even count), 11 -> + 8."""
    s, d = sys256
    rhs, unit = _rhs(s, 300 + per_cta)
    iters = 30
    xr, _ = s.pcg(rhs, iters, 0.0)
    d.pcg_set_dense(False)
    d.pcg_set_resident(False)
    xw, nw = d.pcg_solve(rhs, iters, 0.0)
    trw = d.pcg_trace().copy()
    assert d.pcg_last_kernel() == 0
    tiles = d.pcg_active_cells() // (16 * 128)
    limit = -(-tiles // per_cta)
    assert 4 * limit < tiles <= 12 * limit, (tiles, limit)
    d.pcg_set_grid_limit(limit)
    try:
        d.pcg_set_resident(True)
        xp, n = d.pcg_solve(rhs, iters, 0.0)
        trp = d.pcg_trace().copy()
        assert d.pcg_last_kernel() == 2
        d.pcg_set_resident(2)  # no paging: the list does not fit, the streaming kernel takes it
        x2, n2 = d.pcg_solve(rhs, iters, 0.0)
        assert d.pcg_last_kernel() == 0
        d.pcg_set_resident(True)
        assert n == n2 == nw == iters
        assert H.rel_l2(xp, xr) < TOL, H.rel_l2(xp, xr)
        assert H.rel_l2(xp, xw) < 1e-11
        assert np.allclose(trp[:iters], trw[:iters], rtol=1e-9, atol=0)
        assert not xp[~unit & (rhs == 0)].any()
        # run to run: bit-identical
        xq, _ = d.pcg_solve(rhs, iters, 0.0)
        assert np.array_equal(xq, xp)
        # convergence exit with a prefetched box in flight, zero rhs, reuse
        tol = float(trp[12, 3]) * 1.0000001
        first = int(np.argmax(trp[:, 3] <= tol))
        xc, nc = d.pcg_solve(rhs, iters, tol)
        assert nc == first and d.pcg_last_kernel() == 2
        xrc, nrc = s.pcg(rhs, first + 1, 0.0)
        assert H.rel_l2(xc, xrc) < TOL
        x0, n0 = d.pcg_solve(np.zeros(s.N), iters, 0.0)
        assert n0 == 0 and not x0.any()
        x1, n1 = d.pcg_solve(rhs, 1, 0.0)
        xr1, _ = s.pcg(rhs, 1, 0.0)
        assert n1 == 1 and H.rel_l2(x1, xr1) < TOL
    finally:
        d.pcg_set_grid_limit(0)
        d.pcg_set_resident(True)


@pytest.fixture(scope="module")
def sys1024(ref_mod, scene_dir):
    scene = scenes.dam_break(1024, "flip")
    s = _system(ref_mod, scene_dir, scene, "mt1024", 1.0 / 300.0)
    d = _device(s, scene)
    yield s, d
    d.close()
    s.close()


@pytest.fixture(scope="module")
def sys2048(ref_mod, scene_dir):
    scene = scenes.smoke_test(2048, ppc=4, parameter_handling="grid")
    s = _system(ref_mod, scene_dir, scene, "mt2048", 1.0 / 300.0)
    d = _device(s, scene)
    yield s, d
    d.close()
    s.close()


@pytest.mark.parametrize("limit", [0, 37])
def test_flip_1024_against_reference(sys1024, limit):
    """BASELINE config 2 size: 1024^2 dam break, 512 tiles. Full grid (296 CTAs: 1 - 2 tiles each) and 37 CTAs
    (13 - 14 tiles each); dense and active walk; whole-solve and stepwise."""
    s, d = sys1024
    rhs, unit = _rhs(s, 1024)
    iters = 40
    xr, itr = s.pcg(rhs, iters, 0.0)
    assert itr == iters
    d.pcg_set_grid_limit(limit)
    for dense in (True, False):
        d.pcg_set_dense(dense)
        d.pcg_set_stepwise(False)
        d.pcg_set_resident(False)
        xw, nw = d.pcg_solve(rhs, iters, 0.0)
        d.pcg_set_stepwise(True)
        xs, ns = d.pcg_solve(rhs, iters, 0.0)
        d.pcg_set_stepwise(False)
        d.pcg_set_resident(True)
        assert nw == ns == iters
        assert np.array_equal(xw, xs)
        assert H.rel_l2(xw, xr) < TOL, (dense, H.rel_l2(xw, xr))
        if not dense:
            assert d.pcg_active_cells() < s.N // 4
            # ~60 active tiles: with the full grid the resident kernel holds one tile per CTA, with 37 CTAs two
            xres, nres = d.pcg_solve(rhs, iters, 0.0)
            assert nres == iters and H.rel_l2(xres, xr) < TOL, H.rel_l2(xres, xr)
            assert H.rel_l2(xres, xw) < 1e-11
    if limit:
        # ~60 active tiles over 10 CTAs: 3 resident + 3 paged tiles each
        d.pcg_set_dense(False)
        d.pcg_set_grid_limit(10)
        xpg, npg = d.pcg_solve(rhs, iters, 0.0)
        assert d.pcg_last_kernel() == 2, (d.pcg_last_kernel(), d.pcg_active_cells() // 2048)
        assert npg == iters and H.rel_l2(xpg, xr) < TOL and H.rel_l2(xpg, xw) < 1e-11
        d.pcg_set_grid_limit(limit)
    # the real right-hand side of the scene (hydrostatic column at rest): body forces, then calcPressureRhs
    s.stage("BODY_FORCES")
    rhs2 = s.pressure_rhs()
    xr2, _ = s.pcg(rhs2, iters, 0.0)
    d.pcg_set_dense(False)
    x2, n2 = d.pcg_solve(rhs2, iters, 0.0)
    assert n2 == iters and H.rel_l2(x2, xr2) < TOL
    d.pcg_set_grid_limit(0)


@pytest.mark.parametrize("limit", [0, 64])
def test_smoke_2048_rows_against_reference(sys2048, limit):
    """BASELINE config 3 size: the smoke solver makes every non-solid cell a pressure DOF (flipsmokesolver.cpp:354-444),
    so at 2048^2 (2048 tiles) the dense and the active walk both give every CTA 7 (296 CTAs) or 32 (64 CTAs) tiles."""
    s, d = sys2048
    rhs, unit = _rhs(s, 2048, lone=False)
    assert unit.sum() > 0.7 * s.N
    iters = 20
    xr, itr = s.pcg(rhs, iters, 0.0)
    assert itr == iters
    d.pcg_set_grid_limit(limit)
    for dense in (True, False):
        d.pcg_set_dense(dense)
        xw, nw = d.pcg_solve(rhs, iters, 0.0)
        assert nw == iters
        assert H.rel_l2(xw, xr) < TOL, (dense, H.rel_l2(xw, xr))
    d.pcg_set_stepwise(True)
    xs, ns = d.pcg_solve(rhs, iters, 0.0)
    assert ns == iters and np.array_equal(xs, xw)
    d.pcg_set_stepwise(False)
    d.pcg_set_dense(False)
    d.pcg_set_grid_limit(0)


@pytest.mark.parametrize("world,limit", [(2, 3), (4, 2), (2, 0)])
@pytest.mark.parametrize("dense", [True, False, None, "mixed"], ids=["dense", "active", "resident", "mixed"])
def test_slab_solve_with_many_tiles_per_cta(ref_mod, scene_dir, world, limit, dense):
    """pcgSolveKernel<MG = true> (row slabs; here the ranks share one GPU) with 3 - 8 tiles per CTA: the halo rows
    pushed into the neighbour, the machine-wide barrier and the pipelined walk together, against the reference."""
    scene = scenes.dam_break(256, "flip")
    scene["settings"]["density"] = 0.02
    s = _system(ref_mod, scene_dir, scene, "mtslab", 1.0 / 60.0)
    rhs, unit = _rhs(s, 9)
    iters = 25
    xr, _ = s.pcg(rhs, iters, 0.0)
    mat = s.grid("MATERIAL")
    devs = []
    for r in range(world):
        d = H.make_device(s, scene)
        d.slab_configure(r, world, device_share=world)
        devs.append(d)
    capi.connect_slabs(devs)
    for d in devs:
        d.upload("MATERIAL", mat)
        d.set_step_dt(1.0 / 60.0)
        d.stage("build_matrix")
        # None: active walk through pcgResidentKernel<MG> (1 - 4 resident tiles per CTA; halo rows travel as LL words);
        # "mixed": even ranks resident, odd ranks streaming -- the kernels interoperate (plain halo rows + fence)
        resident = dense is None or (dense == "mixed" and d.rank % 2 == 0)
        d.pcg_set_dense(dense is True)
        d.pcg_set_resident(resident)
        d.pcg_set_grid_limit(limit)
    J = s.J
    outs = {}
    for stepwise in ((False,) if dense in (None, "mixed") else (False, True)):
        for d in devs:
            d.pcg_set_stepwise(stepwise)
        res = capi.run_ranks([lambda d=d: d.pcg_solve(rhs, iters, 0.0) for d in devs])
        x = np.zeros(s.N)
        for d, (xd, nd) in zip(devs, res):
            lo, hi, _ = d.slab_rows()
            x[lo * J: hi * J] = xd[lo * J: hi * J]
            assert nd == iters
        outs[stepwise] = x
        assert H.rel_l2(x, xr) < TOL, (stepwise, H.rel_l2(x, xr))
    if dense in (True, False):
        assert np.array_equal(outs[False], outs[True])
    for d in devs:
        d.close()
    s.close()
