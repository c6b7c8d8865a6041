"""Row-slab decomposition: several ranks -- here several handles sharing ONE GPU, each on its own
host thread, talking through the same peer-memory mailboxes a multi-GPU run uses -- must reproduce the
single-handle result: integer grids exactly, floating-point fields to the rounding of the regrouped
dot-product partials (the only arithmetic that depends on the decomposition)."""
import numpy as np
import pytest

import helpers as H
from flipsolver2d_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def _scene(res, falling=False):
    sc = scenes.dam_break(res, "flip")
    sc["settings"]["density"] = 0.02
    if falling:
        # a block that starts above the floor and straddles the middle of the tank: it falls (gravity is +i) across the
        # slab boundaries of a 2- and a 4-rank split, so particles migrate and ghosts are exchanged
        sc["solver"]["objects"][-1]["verts"] = [[15, 3], [15, 13], [40, 13], [40, 3]]
    return sc


def _slab_devices(s, scene, world, row_bounds=None, **over):
    devs = []
    for r in range(world):
        d = H.make_device(s, scene, **over)
        d.slab_configure(r, world, device_share=world, row_bounds=row_bounds)
        devs.append(d)
    capi.connect_slabs(devs)
    return devs


def _sync_slab(s, d):
    for g in H.STATE_GRIDS:
        d.upload(g, s.grid(g))
    pos, vel, props, bins = s.particles()
    lo, hi, _ = d.slab_rows()
    own = (np.floor(pos[:, 0]) >= lo) & (np.floor(pos[:, 0]) < hi)
    d.upload_particles(pos[own], vel[own], props[:, own])
    if own.any():
        d.set_storage_bins(bins[own])
    d.set_step_dt(s.params()["stepDt"])
    return int(own.sum())


def _assemble(devs, name, J, per_row):
    """Rows every rank owns, stitched together (per_row = elements per grid row of this array)."""
    out = None
    for d in devs:
        a = d.download(name)
        if out is None:
            out = np.zeros_like(a)
        lo, hi, _ = d.slab_rows()
        hi_e = hi
        if name in ("U", "U_VALID", "SAVED_U") and d.rank == d.world - 1:
            hi_e = hi + 1  # the extra U row belongs to the last slab
        out[lo * per_row: hi_e * per_row] = a[lo * per_row: hi_e * per_row]
    return out


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("dense", [False, True])
def test_slab_pcg_matches_single_handle(ref_mod, scene_dir, world, dense):
    scene = _scene(128)
    s = H.make_ref(ref_mod, scene, scene_dir / "slabpcg.json")
    s.stage("FIRST_FRAME_INIT")
    s.set_step_dt(1.0 / 60.0)
    mat = s.grid("MATERIAL")
    rng = np.random.default_rng(7)

    def prep(d):
        d.upload("MATERIAL", mat)
        d.set_step_dt(1.0 / 60.0)
        d.stage("build_matrix")
        d.pcg_set_dense(dense)

    single = H.make_device(s, scene)
    prep(single)
    unit = single.matrix()["is_unit"].astype(bool)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    cases = [(40, 0.0), (200, 1e-6)]
    want = [single.pcg_solve(rhs, it, tol) for it, tol in cases]
    trace1 = single.pcg_trace()
    devs = _slab_devices(s, scene, world)
    for d in devs:
        prep(d)
    J = s.J
    for (it, tol), (x1, n1) in zip(cases, want):
        res = capi.run_ranks([lambda d=d: d.pcg_solve(rhs, it, tol) for d in devs])
        x = np.zeros_like(x1)
        for d, (xr, nr) in zip(devs, res):
            lo, hi, _ = d.slab_rows()
            x[lo * J: hi * J] = xr[lo * J: hi * J]
            assert nr == n1, (nr, n1)
        assert H.rel_l2(x, x1) < 1e-9, H.rel_l2(x, x1)
    tr = devs[0].pcg_trace()
    assert len(tr) == len(trace1)
    assert np.allclose(tr, trace1, rtol=1e-7, atol=0)
    for d in devs:
        d.close()
    single.close()


@pytest.mark.parametrize("world,bounds", [(2, None), (4, None), (2, (0, 48, 128)), (2, (0, 96, 128)), (3, (0, 32, 80, 128))])
def test_slab_substeps_match_single_handle(ref_mod, scene_dir, world, bounds):
    scene = _scene(128, falling=True)
    s = H.make_ref(ref_mod, scene, scene_dir / "slabstep.json")
    s.stage("FIRST_FRAME_INIT")
    s.bump_frame()
    dt = 1.0 / 60.0
    steps = 6
    single = H.make_device(s, scene)
    H.sync_state(s, single)
    for _ in range(steps):
        single.substep(dt)
    devs = _slab_devices(s, scene, world, row_bounds=bounds)  # None: equal row counts; else work-balanced style cuts
    assert sum(_sync_slab(s, d) for d in devs) == s.particle_count()

    def run(d):
        its = []
        for _ in range(steps):
            its.append(d.substep(dt)[1].copy())
        return its

    iters = capi.run_ranks([lambda d=d: run(d) for d in devs])
    for r in range(1, world):
        assert all(np.array_equal(a, b) for a, b in zip(iters[0], iters[r]))  # every rank sees the same PCG
    I, J = s.I, s.J
    assert np.array_equal(_assemble(devs, "MATERIAL", J, J), single.download("MATERIAL"))
    assert np.array_equal(_assemble(devs, "COUNTS", J, J), single.download("COUNTS"))
    assert np.array_equal(_assemble(devs, "U_VALID", J, J), single.download("U_VALID"))
    for name, per_row in (("U", J), ("V", J + 1), ("PRESSURE", J)):
        a, b = _assemble(devs, name, J, per_row), single.download(name)
        assert H.rel_l2(a, b) < 1e-7, (name, H.rel_l2(a, b))
    # the level set below the surface is extrapolated lazily with unbounded radius: a plain download refuses it in
    # slab mode, the collective gather brings all rows to every rank and runs the deferred pass there
    with pytest.raises(capi.Fs2dError):
        devs[0].download("FLUID_SDF")
    capi.run_ranks([lambda d=d: d.slab_gather("FLUID_SDF") for d in devs])
    want_sdf = single.download("FLUID_SDF")
    for d in devs:
        assert np.array_equal(d.download("FLUID_SDF"), want_sdf)  # a minimum and +-1 steps: order free, bit-exact
    capi.run_ranks([lambda d=d: d.slab_gather("U") for d in devs])
    capi.run_ranks([lambda d=d: d.slab_gather("V") for d in devs])
    assert np.array_equal(devs[0].download("U"), _assemble(devs, "U", J, J))
    assert np.array_equal(devs[world - 1].download("V"), _assemble(devs, "V", J, J + 1))
    assert sum(d.particle_count() for d in devs) == single.particle_count()
    parts = [d.download_particles() for d in devs]
    pos = np.concatenate([p[0] for p in parts])
    vel = np.concatenate([p[1] for p in parts])
    p1, v1, _ = single.download_particles()
    # ranks hold consecutive row ranges and each is sorted by cell -> the concatenation is the global order
    assert pos.shape == p1.shape
    assert np.max(np.abs(pos - p1)) < 1e-5
    assert H.rel_l2(vel, v1) < 1e-6
    moved = sum(1 for d, p in zip(devs, parts) if len(p[0]))
    assert moved >= 2  # the fluid really spans more than one slab
    for d in devs:
        d.close()
    single.close()


@pytest.mark.parametrize("kind", ["falling", "source_sink"])
def test_host_solver_slabs_match_single_solver(scene_dir, kind):
    """The C++ host mirror (JsonSceneReader -> FlipSolver::stepFrame) with one solver per slab: seed partition, CFL
    all-gather, reseed stream stitched over the ranks, collective grid accessors."""
    from flipsolver2d_b200 import host_api
    world = 2
    if kind == "falling":
        scene = _scene(128, falling=True)
    else:
        scene = scenes.source_sink(96, "flip")  # emitter (host RNG reseeding every substep) + sink + sloped solid
        scene["settings"]["density"] = 0.02
    path = scenes.write_scene(scene, str(scene_dir / ("hostslab_%s.json" % kind)))
    frames = 3
    single = host_api.Solver(path, quiet=True)
    for _ in range(frames):
        single.step_frame()
    solvers = [host_api.Solver(path, quiet=True, slab=(r, world, world)) for r in range(world)]
    host_api.connect_slabs(solvers)

    def run(s):
        for _ in range(frames):
            s.step_frame()
        return s.stats()

    stats = capi.run_ranks([lambda s=s: run(s) for s in solvers])
    want = single.stats()
    for st in stats:
        assert st["substeps"] == want["substeps"]
        assert st["pressure_iters"] == want["pressure_iters"] and st["density_iters"] == want["density_iters"]
    mats = capi.run_ranks([lambda s=s: s.material() for s in solvers])  # collective accessor: gathers all rows
    for m in mats:
        assert np.array_equal(m, single.material())
    counts = capi.run_ranks([lambda s=s: s.global_particle_count() for s in solvers])
    assert counts == [single.particle_count()] * world
    assert sum(s.particle_count() for s in solvers) == single.particle_count()
    assert min(s.particle_count() for s in solvers) > 0
    for s in solvers:
        s.close()
    single.close()


@pytest.mark.parametrize("sim,handling", [("smoke", "particle"), ("smoke", "grid"), ("fire", "grid"), ("fire", "particle")])
def test_smoke_fire_slabs_match_single_solver(scene_dir, sim, handling):
    """FlipSmokeSolver / FlipFireSolver over row slabs: all non-solid cells are pressure unknowns (dense PCG walk across
    the slab boundary), temperature / soot / fuel grids need halo rows for the buoyancy force, the semi-Lagrangian step
    (grid mode) and the reseeding at the emitter, particles carry two or three property columns across the boundary."""
    from flipsolver2d_b200 import host_api
    world = 2
    scene = scenes.smoke_test(96, parameter_handling=handling, sim_type=sim)
    # the emitter straddles the slab boundary (row 48 of 96 = 25 domain units): reseeding, buoyancy and the particle
    # exchange all happen on both sides of it from the first substep on
    src = [o for o in scene["solver"]["objects"] if o["type"] == "source"][0]
    src["verts"] = [[21, 20], [21, 30], [29, 30], [29, 20]]
    path = scenes.write_scene(scene, str(scene_dir / ("hostslab_%s_%s.json" % (sim, handling))))
    frames = 8
    single = host_api.Solver(path, quiet=True)
    for _ in range(frames):
        single.step_frame()
    solvers = [host_api.Solver(path, quiet=True, slab=(r, world, world)) for r in range(world)]
    host_api.connect_slabs(solvers)

    def run(s):
        for _ in range(frames):
            s.step_frame()
        return s.stats()

    stats = capi.run_ranks([lambda s=s: run(s) for s in solvers])
    want = single.stats()
    for st in stats:
        assert st["substeps"] == want["substeps"]
        assert st["pressure_iters"] == want["pressure_iters"] and st["density_iters"] == want["density_iters"]
    mats = capi.run_ranks([lambda s=s: s.material() for s in solvers])
    for m in mats:
        assert np.array_equal(m, single.material())
    counts = capi.run_ranks([lambda s=s: s.global_particle_count() for s in solvers])
    assert counts == [single.particle_count()] * world
    assert min(s.particle_count() for s in solvers) > 0
    K = 4 if sim == "fire" else 3   # property columns: viscosity, concentration, temperature (+ fuel)
    ds, dd = single.device(K), [s.device(K) for s in solvers]
    for g in ("TEMPERATURE", "CONCENTRATION", "U", "V"):
        capi.run_ranks([lambda d=d: d.slab_gather(g) for d in dd])
        a, b = dd[0].download(g), ds.download(g)
        assert np.array_equal(dd[1].download(g), a), g
        assert H.rel_l2(a, b) < 1e-5, (g, H.rel_l2(a, b))
    for s in solvers:
        s.close()
    single.close()


def test_slab_allgather_and_errors(ref_mod, scene_dir):
    scene = _scene(128)
    s = H.make_ref(ref_mod, scene, scene_dir / "slabmisc.json")
    s.stage("FIRST_FRAME_INIT")
    devs = _slab_devices(s, scene, 2)
    got = capi.run_ranks([lambda d=d: d.slab_allgather([10 + d.rank, -d.rank]) for d in devs])
    for g in got:
        assert g[:, 0].tolist() == [10, 11] and g[:, 1].tolist() == [0, -1]
    # a slab thinner than the halo, or a second configure, is refused
    d = H.make_device(s, scene)
    with pytest.raises(capi.Fs2dError):
        d.slab_configure(0, 8)
    d.slab_configure(0, 1)
    with pytest.raises(capi.Fs2dError):
        d.slab_configure(0, 1)
    d.close()
    for d in devs:
        d.close()


def test_withheld_rank_is_reported_not_ignored(ref_mod, scene_dir):
    """A peer that never shows up: the kernels of the waiting rank give up after their spin limit (~4 s) and flag
    SlabMail::error; the host must turn that into FS2D_ERR_COMM at its next synchronisation point instead of stepping
    on with stale halo rows, and keep reporting it (this rank's halos are stale for good)."""
    scene = _scene(128)
    s = H.make_ref(ref_mod, scene, scene_dir / "slablost.json")
    s.stage("FIRST_FRAME_INIT")
    s.set_step_dt(1.0 / 60.0)
    devs = _slab_devices(s, scene, 2)
    for d in devs:
        d.upload("MATERIAL", s.grid("MATERIAL"))
        d.set_step_dt(1.0 / 60.0)
        d.stage("build_matrix")
    rng = np.random.default_rng(1)
    unit = devs[0].matrix()["is_unit"].astype(bool)
    rhs = np.where(unit, rng.standard_normal(s.N), 0.0)
    with pytest.raises(capi.Fs2dError, match="communication"):
        devs[0].pcg_solve(rhs, 10, 0.0)      # rank 1 never calls
    with pytest.raises(capi.Fs2dError, match="communication"):
        devs[0].slab_allgather([1, 2])
    for d in devs:
        d.close()
    s.close()
