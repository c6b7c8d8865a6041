"""State dump / restore: a checkpoint written in the middle of a run, loaded into a fresh solver of the
same scene, continues bit-identically to the uninterrupted run -- particles (order included), velocities, materials,
level set, smoke grids, frame / substep counters and the mt19937 stream of the reseeding."""
import numpy as np
import pytest

from flipsolver2d_b200 import host_api, scenes

pytestmark = pytest.mark.gpu


def _scene(kind, res):
    if kind == "smoke":
        return scenes.smoke_test(res, ppc=4, parameter_handling="grid"), 3
    if kind == "nbflip":
        return scenes.dam_break(res, "nbflip", viscosity_enabled=True), 2
    return scenes.dam_break(res, "flip"), 2


def _state(solver, props, kind):
    d = solver.device(num_properties=props)
    pos, vel, pr = d.download_particles()
    out = {"pos": pos, "vel": vel, "props": pr, "frame": solver.frame_number()}
    for g in ("U", "V", "MATERIAL", "FLUID_SDF", "VISCOSITY", "COUNTS") + (("TEMPERATURE", "CONCENTRATION") if kind == "smoke" else ()):
        out[g] = d.download(g)
    return out


@pytest.mark.parametrize("kind,res,before,after", [("flip", 128, 7, 9), ("nbflip", 128, 5, 8), ("smoke", 96, 6, 8)])
def test_checkpoint_continues_bit_identically(scene_dir, tmp_path, kind, res, before, after):
    scene, props = _scene(kind, res)
    path = scenes.write_scene(scene, str(scene_dir / ("state_%s.json" % kind)))
    a = host_api.Solver(path, quiet=True)
    for _ in range(before):
        a.step_substep()
    ck = str(tmp_path / "state.bin")
    a.save_state(ck)
    for _ in range(after):
        a.step_substep()
    ref = _state(a, props, kind)
    a.close()

    b = host_api.Solver(path, quiet=True)
    b.load_state(ck)
    for _ in range(after):
        b.step_substep()
    got = _state(b, props, kind)
    b.close()
    assert got["frame"] == ref["frame"]
    for k, v in ref.items():
        if k == "frame":
            continue
        assert np.array_equal(np.asarray(v), np.asarray(got[k])), k
    assert ref["pos"].shape[0] > 0


def test_state_blob_is_rejected_by_a_different_handle(scene_dir):
    p1 = scenes.write_scene(scenes.dam_break(64, "flip"), str(scene_dir / "state_a.json"))
    p2 = scenes.write_scene(scenes.dam_break(96, "flip"), str(scene_dir / "state_b.json"))
    a, b = host_api.Solver(p1, quiet=True), host_api.Solver(p2, quiet=True)
    a.prepare()
    b.prepare()
    blob = a.device(num_properties=2).state_save()
    with pytest.raises(Exception):
        b.device(num_properties=2).state_load(blob)
    # and the writer takes its own blob back
    a.device(num_properties=2).state_load(blob)
    a.step_substep()
    a.close()
    b.close()
