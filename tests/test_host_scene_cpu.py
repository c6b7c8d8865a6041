"""Frame-0 parity of the C++ host mirror (JsonSceneReader + FlipSolver::firstFrameInit's host half)
against the oracle: scene rasterisation (materials, solid sdf, ids) and the mt19937-seeded particles,
bit-exact. CPU only -- the host half needs no GPU."""
import numpy as np
import pytest

import helpers as H
from flipsolver2d_b200 import host_api, scenes


def _nonsquare():
    sc = scenes.source_sink(48, "flip")
    sc["settings"]["domainSizeI"] = 50
    sc["settings"]["domainSizeJ"] = 35
    return sc


CASES = {
    "dam64": lambda: scenes.dam_break(64, "flip"),
    "dam100_ppc4": lambda: scenes.dam_break(100, "flip", ppc=4, seed=7),
    "src48": lambda: scenes.source_sink(48, "flip"),
    "nonsquare": _nonsquare,
    "smoke64": lambda: scenes.smoke_test(64),
    "fire48": lambda: scenes.smoke_test(48, sim_type="fire"),
    "nbflip64": lambda: scenes.dam_break(64, "nbflip"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_frame0_matches_reference(ref_mod, scene_dir, name):
    scene = CASES[name]()
    path = scene_dir / ("host_%s.json" % name)
    s = H.make_ref(ref_mod, scene, path)
    s.stage("FIRST_FRAME_INIT")
    h = host_api.Solver(str(path))
    h.prepare_host()
    assert (h.I, h.J) == (s.I, s.J)
    for g in ("MATERIAL", "SOLID_SDF", "SOLID_ID", "EMITTER_ID", "DIVERGENCE_CONTROL", "VISCOSITY"):
        assert np.array_equal(h.host_grid(g), s.grid(g)), g
    if scene["settings"]["simType"] == "nbflip":
        assert np.array_equal(h.host_grid("FLUID_SDF"), s.grid("FLUID_SDF"))
    rp, rv, rprops, _ = s.particles()
    hp, hv, hprops = h.seed_particles(s.property_count())
    assert len(hp) == len(rp)
    if len(rp):
        rp, rv, rprops = H.canonical(rp, rv, rprops, J=s.J)
        hp, hv, hprops = H.canonical(hp, hv, hprops, J=s.J)
        assert np.array_equal(hp, rp)
        assert np.array_equal(hv, rv)
        assert np.array_equal(hprops, rprops)
    h.close()
    s.close()


def test_loadjson_returns_null_on_bad_scene(tmp_path):
    bad = tmp_path / "bad.json"
    bad.write_text("{ not json")
    with pytest.raises(RuntimeError):
        host_api.Solver(str(bad))
    with pytest.raises(RuntimeError):
        host_api.Solver(str(tmp_path / "missing.json"))
