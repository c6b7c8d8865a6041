"""Stage-wise parity for the stages that only the frame-level tests used to reach: the full density correction
(flipsolver2d.cpp:164-186,261-303), reseedParticles (:627-680) with the host mt19937 stream, the narrow-band solver's
semi-Lagrangian grids and combine passes (nbflipsolver.cpp:66-109,213-225,366-427), smoke buoyancy / decay / centred
parameters (flipsmokesolver.cpp:23-52,104-130,509-558) and fire combustion (flipfiresolver.cpp:35-106).

The device is built by the host mirror (JsonSceneReader -> frame-0 set-up: static tables, source level set), advanced
as many frames as the reference, then OVERWRITTEN with the reference's state, so that both sides run one stage from
identical inputs. Oracle = the unmodified reference (strict build, one ThreadPool worker)."""
import numpy as np
import pytest

import helpers as H
from flipsolver2d_b200 import host_api, scenes

pytestmark = pytest.mark.gpu

SUM_TOL = 2e-6   # float sums accumulated in a different order (P2G-type gathers)


def _pair(ref_mod, scene_dir, scene, name, frames, props):
    path = scene_dir / (name + ".json")
    s = H.make_ref(ref_mod, scene, path)
    h = host_api.Solver(str(path), convergence_threads=s.threads)
    if frames == 0:
        s.stage("FIRST_FRAME_INIT")
        s.bump_frame()
        h.prepare()
    for _ in range(frames):
        s.step_frame()
        h.step_frame()
    d = h.device(num_properties=props)
    sim = scene["settings"]["simType"]
    H.sync_state(s, d, sim)
    return s, h, d


def _particles(s, d, exact=True, tol=0.0):
    rp, rv, rprops, _ = s.particles()
    dp, dv, dprops = d.download_particles()
    assert len(rp) == len(dp), (len(rp), len(dp))
    rp, rv, rprops = H.canonical(rp, rv, rprops, J=s.J)
    dp, dv, dprops = H.canonical(dp, dv, dprops, J=s.J)
    if exact:
        assert np.array_equal(rp, dp)
        assert np.array_equal(rv, dv)
        assert np.array_equal(rprops, dprops)
    else:
        assert H.max_abs(dp, rp) <= tol, H.max_abs(dp, rp)
    return rp, dp, rprops, dprops


def test_density_correction_moves_particles_like_the_reference(ref_mod, scene_dir):
    """densityCorrection in the regime where its solve converges (at BASELINE's density 0.5 it runs into the iteration cap
    and returns WITHOUT moving a particle, flipsolver2d.cpp:179-182 -- that branch is covered by the 1024^2 sweep):
    density grid, rhs, PCG with the reference's convergence test, then adjustParticlesByDensity."""
    scene = scenes.dam_break(64, "flip")
    scene["settings"]["density"] = 0.02
    s, h, d = _pair(ref_mod, scene_dir, scene, "var_density", 3, 2)
    s.set_step_dt(1.0 / 120.0)
    d.set_step_dt(1.0 / 120.0)
    s.stage("BUILD_MATRIX")
    d.stage("build_matrix")
    p0 = s.particles()[0].copy()
    s.stage("UPDATE_DENSITY_GRID")
    _, it_ref = s.pcg(s.density_rhs(), int(s.params()["pcgIterLimit"]), s.params()["projectTolerance"])
    assert it_ref < int(s.params()["pcgIterLimit"])
    s.stage("DENSITY_CORRECTION")
    it = d.stage_iters("density_correction")
    assert it == it_ref, (it, it_ref)
    p1 = s.particles()[0]
    moved = np.abs(H.canonical(p1, J=s.J) - H.canonical(p0, J=s.J)).max()
    assert moved > 1e-4, moved   # the stage really pushes particles
    # pos += grad(p) * dt^2 / (rho dx^2): the pressure agrees to ~1e-7 relative (regrouped dot products), the push is a
    # small correction on positions of magnitude <= 64 -> 2e-5 cell units absolute
    _particles(s, d, exact=False, tol=2e-5)
    assert np.array_equal(np.sort(d.storage_bins()), np.sort(s.particles()[3]))  # nobody is re-filed (flipsolver2d.cpp:427)
    h.close()
    s.close()


def _std_uniform_floats(raw_u32):
    """libstdc++ std::uniform_real_distribution<float>(0, 1) on std::mt19937: generate_canonical<float, 24> takes ONE
    32-bit draw, converts it to float (round to nearest) and divides by 2^32; a result of 1.0 becomes nextafter(1, 0)."""
    u = raw_u32.astype(np.float32) / np.float32(4294967296.0)
    return np.where(u >= np.float32(1.0), np.nextafter(np.float32(1.0), np.float32(0.0)), u).astype(np.float32)


def test_reseed_stage_with_the_host_stream(ref_mod, scene_dir):
    """countParticles + reseedParticles right after frame-0 seeding: the SOURCE cells are empty, so every one receives
    particlesPerCell / 2 jittered particles drawn from the solver's std::mt19937 (seed 0) where seedInitialFluid left
    it: 2 draws per seeded particle (flipsolver2d.cpp:686-706,1013-1019). The device plans the candidates, the host draws
    -- here numpy's MT19937 with the same init_genrand seeding -- and the device places them."""
    scene = scenes.source_sink(64, "flip")
    s, h, d = _pair(ref_mod, scene_dir, scene, "var_reseed", 0, 2)
    seeded = s.particle_count()
    s.stage("COUNT_PARTICLES")
    d.stage("count_particles")
    assert np.array_equal(s.grid("COUNTS"), d.download("COUNTS"))
    s.stage("RESEED")
    added = s.particle_count() - seeded
    assert added > 0
    bitgen = np.random.MT19937()
    bitgen._legacy_seeding(int(scene["settings"]["seed"]))
    raw = bitgen.random_raw(2 * seeded + 2 * added + 64).astype(np.uint32)
    drawn = []

    def uniforms(n):
        drawn.append(n)
        return _std_uniform_floats(raw[2 * seeded: 2 * seeded + 2 * n])

    planned = d.reseed(uniforms)
    assert planned == added and drawn == [added]
    assert d.particle_count() == s.particle_count()
    _particles(s, d, exact=True)   # positions (i + u, j + u'), source velocity / viscosity: bit-exact
    h.close()
    s.close()


@pytest.mark.parametrize("viscous", [False, True])
def test_nbflip_stage_sequence(ref_mod, scene_dir, viscous):
    """NBFlipSolver::step() up to the pressure solve, stage by stage at BASELINE's density 0.5: RK4 advection +
    pruneNarrowBand + semi-Lagrangian U / V / sdf / viscosity (advect), P2G, gridUpdate (updateSdf,
    extrapolateLevelsetOutside, afterTransfer with combineAdvectedGrids / combineLevelset, extrapolateLevelsetInside),
    materials, body forces, matrix, rhs."""
    scene = scenes.dam_break(64, "nbflip", viscosity_enabled=viscous)
    s, h, d = _pair(ref_mod, scene_dir, scene, "var_nbflip_%d" % viscous, 2, 2)
    d.set_sdf_band(0)   # the reference's unbounded level-set walks: the whole field is compared bit for bit below
    dt = 1.0 / 120.0
    s.set_step_dt(dt)
    d.set_step_dt(dt)
    s.stage("ADVECT")
    s.stage("PRUNE_REBIN")
    d.stage("advect")
    d.stage("nbflip_advect_grids")
    d.stage("sort_particles")
    assert d.particle_count() == s.particle_count()
    _particles(s, d, exact=True)
    s.stage("P2G")
    d.stage("particle_to_grid")
    for g in ("U", "V", "VISCOSITY"):
        assert H.rel_l2(d.download(g), s.grid(g)) < SUM_TOL, g
    # continue from identical transfers (the advected grids computed above stay on the device)
    for g in ("U", "V", "U_VALID", "V_VALID", "VISCOSITY", "KNOWN_CENTERED", "TEST", "DIVERGENCE_CONTROL"):
        d.upload(g, s.grid(g))
    s.stage("EXTRAPOLATE_VEL")
    s.stage("SAVE_VELOCITY")
    d.stage("extrapolate_velocity", 10)
    d.stage("save_velocity")
    for g in ("U", "V", "U_VALID", "V_VALID", "SAVED_U", "SAVED_V"):
        assert np.array_equal(s.grid(g), d.download(g)), g
    s.stage("GRID_UPDATE")   # NBFlipSolver::gridUpdate (nbflipsolver.cpp:213-225): ... updateMaterials, applyBodyForces
    d.stage("update_sdf")
    d.stage("extrapolate_sdf_outside")
    d.stage("after_transfer")
    d.stage("extrapolate_sdf_inside")
    d.stage("update_materials")
    d.stage("apply_body_forces")
    for g in ("FLUID_SDF", "MATERIAL", "U", "V", "U_VALID", "V_VALID", "VISCOSITY"):
        assert np.array_equal(s.grid(g), d.download(g)), g
    s.stage("BUILD_MATRIX")
    d.stage("build_matrix")
    d.stage("pressure_rhs")
    assert np.array_equal(s.pressure_rhs(), d.download("RHS"))
    if viscous:
        s.stage("VELOCITY_FROM_SOLIDS")
        d.stage("velocity_from_solids")
        s.stage("VISCOSITY")
        it = d.stage_iters("apply_viscosity")
        assert it > 0
        assert H.rel_l2(d.download("U"), s.grid("U")) < 1e-6
        assert H.rel_l2(d.download("V"), s.grid("V")) < 1e-6
    h.close()
    s.close()


@pytest.mark.parametrize("handling", ["particle", "grid"])
def test_smoke_stages(ref_mod, scene_dir, handling):
    """Smoke at stage level: advection (+ semi-Lagrangian temperature / soot in GRID mode), centred parameters to grid,
    buoyancy body force, matrix over all non-solid cells + rhs with sinks / sources, G2P with decay."""
    scene = scenes.smoke_test(64, parameter_handling=handling)
    s, h, d = _pair(ref_mod, scene_dir, scene, "var_smoke_" + handling, 2, 3)
    dt = 1.0 / 120.0
    s.set_step_dt(dt)
    d.set_step_dt(dt)
    s.stage("ADVECT")
    s.stage("PRUNE_REBIN")
    d.stage("advect")
    d.stage("sort_particles")
    assert d.particle_count() == s.particle_count()
    _particles(s, d, exact=True)
    for g in ("TEMPERATURE", "CONCENTRATION"):
        assert np.array_equal(s.grid(g), d.download(g)), g   # GRID mode: eulerAdvectParameters (flipsmokesolver.cpp:211-233)
    s.stage("P2G")
    d.stage("particle_to_grid")
    for g in ("U", "V", "TEMPERATURE", "CONCENTRATION"):
        assert H.rel_l2(d.download(g), s.grid(g)) < SUM_TOL, g
    H.sync_state(s, d, "smoke")
    s.stage("AFTER_TRANSFER")
    d.stage("after_transfer")
    for g in ("U", "V", "TEMPERATURE", "CONCENTRATION", "DIVERGENCE_CONTROL"):
        assert np.array_equal(s.grid(g), d.download(g)), g
    s.stage("EXTRAPOLATE_VEL")
    s.stage("SAVE_VELOCITY")
    s.stage("BODY_FORCES")
    d.stage("extrapolate_velocity", 10)
    d.stage("save_velocity")
    d.stage("apply_body_forces")
    for g in ("U", "V"):
        assert np.array_equal(s.grid(g), d.download(g)), g   # buoyancy (flipsmokesolver.cpp:23-52)
    s.stage("BUILD_MATRIX")
    d.stage("build_matrix")
    rm, dm = s.matrix(), d.matrix()
    for k in ("is_unit", "mask", "count"):
        assert np.array_equal(rm[k], dm[k]), k
    d.stage("pressure_rhs")
    assert np.array_equal(s.pressure_rhs(), d.download("RHS"))   # flipsmokesolver.cpp:73-102
    s.stage("PARTICLE_UPDATE")
    d.stage("particle_update")
    _particles(s, d, exact=True)   # PIC/FLIP blend + temperature / concentration decay (flipsmokesolver.cpp:104-130)
    h.close()
    s.close()


@pytest.mark.parametrize("handling", ["particle", "grid"])
def test_fire_combustion_stage(ref_mod, scene_dir, handling):
    """FlipFireSolver::particleUpdate = smoke update + combustionUpdate (flipfiresolver.cpp:35-106,149-153): fuel above
    the ignition temperature burns into soot, heat and divergence; per particle (PARTICLE mode) or per cell (GRID)."""
    scene = scenes.smoke_test(64, parameter_handling=handling, sim_type="fire")
    s, h, d = _pair(ref_mod, scene_dir, scene, "var_fire_" + handling, 3, 4)
    dt = 1.0 / 120.0
    s.set_step_dt(dt)
    d.set_step_dt(dt)
    fuel0 = s.grid("FUEL").copy()
    pf0 = float(s.particles()[2][3].sum())
    s.stage("PARTICLE_UPDATE")
    d.stage("particle_update")
    _particles(s, d, exact=True)
    for g in ("FUEL", "TEMPERATURE", "CONCENTRATION", "DIVERGENCE_CONTROL"):
        assert np.array_equal(s.grid(g), d.download(g)), g
    burnt = (fuel0 != s.grid("FUEL")).any() or pf0 != float(s.particles()[2][3].sum())
    assert burnt, "the scene never ignites: the combustion branch was not exercised"
    h.close()
    s.close()
