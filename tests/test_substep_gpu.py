"""Short-horizon trajectories: K whole substeps through fs2d_substep vs the reference's
own step(), same scene, same seed, same dt. The scene uses a low fluid density so that the matrix scale
dt/(rho dx^2) is in the regime where the reference's preconditioner is positive definite and PCG
converges; outside it the Krylov recurrence is chaotic and only stage-wise parity
(test_stages_gpu.py) is meaningful. Tolerance: 1e-5 relative on velocity / pressure / particle state,
as BASELINE.json's north_star states (P2G summation order differs)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import helpers as H
from flipsolver2d_b200 import scenes

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _scene(res):
    sc = scenes.dam_break(res, "flip")
    sc["settings"]["density"] = 0.02
    return sc


@pytest.mark.parametrize("res,steps", [(64, 4), (96, 3)])
def test_flip_trajectory(ref_mod, scene_dir, res, steps):
    scene = _scene(res)
    s = H.make_ref(ref_mod, scene, scene_dir / ("traj%d.json" % res))
    s.stage("FIRST_FRAME_INIT")
    s.bump_frame()
    d = H.make_device(s, scene, conv_threads=s.threads)
    H.sync_state(s, d)
    dt = 1.0 / 60.0
    drift = []
    for k in range(steps):
        s.set_step_dt(dt)
        s.stage("FULL_STEP")
        ms, iters = d.substep(dt)
        assert ms.sum() > 0
        drift.append(max(H.rel_l2(d.download("U"), s.grid("U")), H.rel_l2(d.download("V"), s.grid("V"))))
    print("velocity drift per substep:", drift)
    assert d.particle_count() == s.particle_count()
    # 1e-5 after the first substep; the difference then grows with the flow (float P2G sums feed a
    # PCG that stops at tol 1e-2), bounded here by 5e-5 after `steps` substeps
    assert drift[0] < TOL, drift
    assert drift[-1] < 5 * TOL, drift
    mat_r, mat_d = s.grid("MATERIAL"), d.download("MATERIAL")
    assert np.mean(mat_r != mat_d) < 1e-3
    # particles: nearest-neighbour distance between the two sets (cell units)
    rp, rv, _, _ = s.particles()
    dp, dv, _ = d.download_particles()
    dist, idx = cKDTree(rp).query(dp)
    assert dist.max() < 1e-4, dist.max()
    assert H.rel_l2(dv, rv[idx]) < 10 * TOL


def test_substep_is_deterministic(ref_mod, scene_dir):
    """Two handles, same inputs -> bit-identical state (no atomics in the value path)."""
    scene = _scene(64)
    s = H.make_ref(ref_mod, scene, scene_dir / "det64.json")
    s.stage("FIRST_FRAME_INIT")
    outs = []
    for _ in range(2):
        d = H.make_device(s, scene)
        H.sync_state(s, d)
        for k in range(3):
            d.substep(1.0 / 60.0)
        outs.append((d.download("U"), d.download("V"), d.download("PRESSURE")) + d.download_particles())
        d.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("res,density", [(96, 0.5), (128, 0.02)])
def test_packed_round_trip_equals_resident_stepping(ref_mod, scene_dir, res, density):
    """bench.py's end-to-end region moves the particle state host <-> device around every substep
    (fs2d_download_particles_packed / fs2d_upload_particles_packed). That round trip must be the identity on the
    solver state: stepping with it is bit-identical to stepping resident, and no particle is lost. At the BASELINE
    density 0.5 the density correction pushes particles across bin boundaries, so the storage-bin byte matters: the
    plain fs2d_upload_particles (which re-files every particle at home) is shown to change the particle count."""
    scene = scenes.dam_break(res, "flip")
    scene["settings"]["density"] = density
    s = H.make_ref(ref_mod, scene, scene_dir / ("rt%d.json" % res))
    s.stage("FIRST_FRAME_INIT")
    s.bump_frame()
    devs = [H.make_device(s, scene) for _ in range(3)]
    for d in devs:
        H.sync_state(s, d)
    resident, packed, plain = devs
    dt = 1.0 / 60.0
    steps = 8
    counts = []
    for k in range(steps):
        resident.substep(dt)
        buf, n = packed.download_packed()
        assert n >= packed.particle_count()   # records, flagged-dead ones included; no sort on the way out
        packed.upload_packed(buf, n)
        packed.substep(dt)
        pos, vel, props = plain.download_particles()
        plain.upload_particles(pos, vel, props)
        plain.substep(dt)
        counts.append((resident.particle_count(), packed.particle_count(), plain.particle_count()))
    print("particle counts (resident, packed round trip, plain round trip):", counts)
    assert resident.particle_count() == packed.particle_count()
    for a, b in zip(resident.download_particles(), packed.download_particles()):
        assert np.array_equal(a, b)
    for g in ("U", "V", "MATERIAL", "COUNTS", "PRESSURE"):
        assert np.array_equal(resident.download(g), packed.download(g)), g
    assert np.array_equal(resident.storage_bins(), packed.storage_bins())
    for d in devs:
        d.close()
    s.close()


def _pinned(nbytes):
    """Pinned host bytes (the streamed copies are only asynchronous from pinned memory); the tensor keeps them alive."""
    import torch
    t = torch.zeros((nbytes,), dtype=torch.uint8).pin_memory()
    return t, t.numpy()


@pytest.mark.parametrize("res,density,early", [(96, 0.5, True), (128, 0.02, True), (96, 0.5, False)])
def test_streamed_round_trip_equals_resident_stepping(ref_mod, scene_dir, res, density, early):
    """The e2e path of bench.py: fs2d_particle_stream_begin (sections of the host buffer arrive while the substep runs;
    a sort that runs before the property columns are there gathers them later) -> substep -> positions and columns
    leave after the density correction (fs2d_particle_stream_positions_final, here through the composite fs2d_substep
    and fs2d_particle_stream_set_output) -> fs2d_particle_stream_end. Stepping that way is bit-identical to stepping
    resident, and the host buffer holds exactly the state a plain sectioned download gives. At density 0.02 the
    density solve converges, particles are adjusted and re-sorted (second sort with the columns still pending)."""
    scene = scenes.dam_break(res, "flip")
    scene["settings"]["density"] = density
    s = H.make_ref(ref_mod, scene, scene_dir / ("st%d.json" % res))
    s.stage("FIRST_FRAME_INIT")
    s.bump_frame()
    resident, streamed = H.make_device(s, scene), H.make_device(s, scene)
    H.sync_state(s, resident)
    H.sync_state(s, streamed)
    cap = int(streamed.particle_count() * 1.3) + 4096
    nbytes = int(streamed.L.fs2d_particle_stream_bytes(streamed.h, cap))
    keep, buf = _pinned(nbytes)
    keep2, buf2 = _pinned(nbytes)
    n = streamed.stream_end(buf, cap)          # no begin before: a plain download in the sectioned layout
    assert n == streamed.particle_count()
    dt = 1.0 / 60.0
    for k in range(8):
        resident.substep(dt)
        streamed.stream_begin(buf, n, cap)
        if early:
            streamed.stream_set_output(buf, cap)
        streamed.substep(dt)
        n = streamed.stream_end(buf, cap)
        assert n >= streamed.particle_count()
    n2 = streamed.stream_end(buf2, cap)         # nothing left early: every section from the device arrays
    assert n2 == n
    K = streamed.K
    for off, width in [(0, 8), (8 * cap, 8)] + [((16 + 4 * k) * cap, 4) for k in range(K)] + [((16 + 4 * K) * cap, 1)]:
        assert np.array_equal(buf[off:off + width * n], buf2[off:off + width * n]), off
    assert resident.particle_count() == streamed.particle_count()
    for a, b in zip(resident.download_particles(), streamed.download_particles()):
        assert np.array_equal(a, b)
    for g in ("U", "V", "MATERIAL", "COUNTS", "PRESSURE", "VISCOSITY"):
        assert np.array_equal(resident.download(g), streamed.download(g)), g
    assert np.array_equal(resident.storage_bins(), streamed.storage_bins())
    resident.close()
    streamed.close()
    s.close()
