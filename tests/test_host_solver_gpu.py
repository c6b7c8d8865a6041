"""The C++ host mirror end to end on the GPU: JsonSceneReader::loadJson -> stepFrame (CFL sub-stepping,
host mt19937 reseeding, SolverStats) against the reference's own stepFrame on the same scene file."""
import numpy as np
import pytest

import helpers as H
from flipsolver2d_b200 import host_api, scenes

pytestmark = pytest.mark.gpu


def _low_density(scene):
    scene["settings"]["density"] = 0.02  # SPD regime of the reference's preconditioner
    return scene


@pytest.mark.parametrize("name,frames", [("dam64", 3), ("src64", 3)])
def test_step_frame_matches_reference(ref_mod, scene_dir, name, frames):
    scene = _low_density(scenes.dam_break(64, "flip") if name == "dam64" else scenes.source_sink(64, "flip"))
    path = scene_dir / ("hostgpu_%s.json" % name)
    s = H.make_ref(ref_mod, scene, path)
    h = host_api.Solver(str(path), convergence_threads=s.threads)
    for f in range(frames):
        s.step_frame()
        h.step_frame()
        rs, hs = s.stats(), h.stats()
        assert hs["substeps"] == rs["substeps"], (f, hs, rs)
    assert h.frame_number() == s.frame_number() == frames
    d = h.device(num_properties=2)
    assert h.particle_count() == s.particle_count()
    assert H.rel_l2(d.download("U"), s.grid("U")) < 5e-5
    assert H.rel_l2(d.download("V"), s.grid("V")) < 5e-5
    assert np.mean(d.download("MATERIAL") != s.grid("MATERIAL")) < 2e-3
    assert np.array_equal(d.download("SOLID_SDF"), s.grid("SOLID_SDF"))
    # accessors hand out host copies of the device state
    assert np.array_equal(h.material(), d.download("MATERIAL"))
    assert h.bin_sizes().sum() == h.particle_count()
    st = h.stats()
    assert st["frame_ms"] > 0 and st["timings"].sum() > 0
    h.close()
    s.close()


def test_substep_granular_stepping_equals_step_frame(scene_dir):
    scene = _low_density(scenes.dam_break(64, "flip"))
    path = scenes.write_scene(scene, str(scene_dir / "hostgpu_gran.json"))
    a, b = host_api.Solver(path), host_api.Solver(path)
    for _ in range(2):
        a.step_frame()
        while not b.step_substep():
            pass
    da, db = a.device(2), b.device(2)
    for g in ("U", "V", "MATERIAL", "FLUID_SDF"):
        assert np.array_equal(da.download(g), db.download(g)), g
    pa, pb = da.download_particles(), db.download_particles()
    for x, y in zip(pa, pb):
        assert np.array_equal(x, y)
    a.close()
    b.close()


@pytest.mark.parametrize("kind", ["flip", "nbflip_viscous", "smoke", "fire_grid"])
def test_streamed_substeps_equal_resident_substeps(scene_dir, kind):
    """FlipSolver::stepSubstepStreamed (particle state in a pinned host buffer, copies overlapping the stages; the
    scenes have a source, so reseeded records are appended after the positions have left) against stepSubstep. The
    liquid solver sends positions and property columns early; nbflip / smoke / fire keep their own step() and only use
    the streamed upload (sort before the columns arrive, smoke decay rewriting columns) and the download at the end."""
    import torch
    K = 2
    if kind == "flip":
        scene = scenes.source_sink(96, "flip")
    elif kind == "nbflip_viscous":
        scene = scenes.source_sink(96, "nbflip")
        scene["settings"]["viscosityEnabled"] = True
        scene["settings"]["density"] = 0.02
    elif kind == "smoke":
        scene = scenes.smoke_test(96, parameter_handling="particle", sim_type="smoke")
        K = 3   # viscosity, concentration, temperature
    else:
        scene = scenes.smoke_test(96, parameter_handling="grid", sim_type="fire")
        K = 4   # + fuel
    path = scenes.write_scene(scene, str(scene_dir / ("hostgpu_streamed_%s.json" % kind)))
    a, b = host_api.Solver(path), host_api.Solver(path)
    a.prepare()
    b.prepare()
    db = b.device(K)
    cap = int(b.particle_count() * 1.5) + 20000
    pinned = torch.zeros((int(db.L.fs2d_particle_stream_bytes(db.h, cap)),), dtype=torch.uint8).pin_memory()
    n = db.stream_end(pinned.numpy(), cap)
    for _ in range(12):
        fa = a.step_substep()
        fb, n = b.step_substep_streamed(pinned.data_ptr(), cap, n)
        assert fa == fb
    da = a.device(K)
    assert a.particle_count() == b.particle_count() > 0
    grids = ("U", "V", "MATERIAL", "PRESSURE") + (("TEMPERATURE", "CONCENTRATION") if kind in ("smoke", "fire_grid") else ("VISCOSITY",))
    for g in grids:
        assert np.array_equal(da.download(g), db.download(g)), g
    for x, y in zip(da.download_particles(), db.download_particles()):
        assert np.array_equal(x, y)
    a.close()
    b.close()


def test_smoke_scene_steps(ref_mod, scene_dir):
    """smoke_test scene (source + sink + wedge), particle mode: frame loop runs and agrees with the reference."""
    scene = scenes.smoke_test(64)
    path = scene_dir / "hostgpu_smoke.json"
    s = H.make_ref(ref_mod, scene, path)
    h = host_api.Solver(str(path), convergence_threads=s.threads)
    for _ in range(2):
        s.step_frame()
        h.step_frame()
    assert h.particle_count() == s.particle_count()
    d = h.device(num_properties=3)
    assert np.array_equal(d.download("MATERIAL"), s.grid("MATERIAL"))
    h.close()
    s.close()


@pytest.mark.parametrize("sim,visc", [("nbflip", False), ("nbflip", True), ("flip", True), ("flip", "heavy")])
def test_nbflip_and_viscous_frames_match_reference(ref_mod, scene_dir, sim, visc):
    """BASELINE config 4 at test size: narrow-band FLIP (semi-Lagrangian grids, band prune / combine) and the implicit
    viscosity stage with the re-projection that follows it, frame loop against the reference's own stepFrame."""
    scene = _low_density(scenes.dam_break(64, sim, viscosity_enabled=bool(visc)))
    if visc == "heavy":
        scene["settings"]["heavyViscosity"] = True  # HeavyViscosityModel (viscositymodel.cpp:164-470) through JsonSceneReader
    path = scene_dir / ("hostgpu_%s_%s.json" % (sim, visc))
    s = H.make_ref(ref_mod, scene, path)
    h = host_api.Solver(str(path), convergence_threads=s.threads)
    for f in range(3):
        s.step_frame()
        h.step_frame()
        rs, hs = s.stats(), h.stats()
        assert hs["substeps"] == rs["substeps"], (f, hs, rs)
        if visc:
            assert hs["viscosity_iters"] == rs["viscosity_iters"] and rs["viscosity_iters"] > 0, (f, hs, rs)
    d = h.device(num_properties=2)
    assert abs(h.particle_count() - s.particle_count()) <= max(2, s.particle_count() // 500)
    assert H.rel_l2(d.download("U"), s.grid("U")) < 1e-3
    assert H.rel_l2(d.download("V"), s.grid("V")) < 1e-3
    assert np.mean(d.download("MATERIAL") != s.grid("MATERIAL")) < 5e-3
    h.close()
    s.close()


@pytest.mark.parametrize("sim,handling", [("smoke", "grid"), ("smoke", "particle"), ("fire", "particle"), ("fire", "grid")])
def test_smoke_fire_frames_match_reference(ref_mod, scene_dir, sim, handling):
    """BASELINE config 3 at test size: smoke with grid-advected temperature / soot (and the fire subclass)."""
    scene = scenes.smoke_test(64, parameter_handling=handling, sim_type=sim)
    path = scene_dir / ("hostgpu_%s_%s.json" % (sim, handling))
    s = H.make_ref(ref_mod, scene, path)
    h = host_api.Solver(str(path), convergence_threads=s.threads)
    for _ in range(3):
        s.step_frame()
        h.step_frame()
    d = h.device(num_properties=3 if sim == "smoke" else 4)
    assert h.particle_count() == s.particle_count()
    assert np.array_equal(d.download("MATERIAL"), s.grid("MATERIAL"))
    assert H.rel_l2(d.download("U"), s.grid("U")) < 1e-3
    assert H.rel_l2(d.download("TEMPERATURE"), s.grid("TEMPERATURE")) < 1e-3
    assert H.rel_l2(d.download("CONCENTRATION"), s.grid("CONCENTRATION")) < 1e-3
    if sim == "fire":
        # FlipFireSolver::combustionUpdate (flipfiresolver.cpp:35-106): fuel burnt into soot and heat
        assert H.rel_l2(d.download("FUEL"), s.grid("FUEL")) < 1e-3
        pr = s.particles()[2]
        pd = d.download_particles()[2]
        assert abs(float(pr[3].sum()) - float(pd[3].sum())) <= 1e-3 * max(1.0, abs(float(pr[3].sum())))  # fuel column
    h.close()
    s.close()
