"""Host-side logic of the row-slab decomposition with two real processes (gloo, CPU): every rank loads the same
scene, computes the slab boundaries and its share of the seed particles on its own; the ranks must agree on the
boundaries, their shares must be disjoint and complete, and the 256-byte blob exchange bench.py performs
(all_gather of opaque byte buffers) must deliver every rank's blob to every other rank. The data path itself
(peer-mapped device memory) needs GPUs: tests/test_slab_gpu.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, scene_path, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from flipsolver2d_b200 import host_api

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = host_api.Solver(scene_path, quiet=True)
    s.prepare_host()                       # rasterise + seed on the host (same mt19937 stream on every rank)
    bounds = s.slab_bounds(world)
    pos, _, _ = s.seed_particles(2)
    rows = np.floor(pos[:, 0]).astype(np.int64)
    mine = (rows >= bounds[rank]) & (rows < bounds[rank + 1])
    # the ranks agree on the table
    t = torch.from_numpy(bounds.astype(np.int64))
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    # shares are disjoint and complete: the owned counts add up to the seed count
    c = torch.tensor([int(mine.sum())], dtype=torch.int64)
    dist.all_reduce(c)
    # opaque blob exchange as in bench.py
    blob = torch.frombuffer(bytearray(bytes([rank + 1]) * 256), dtype=torch.uint8).clone()
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    blob_ok = all(int(b[0]) == r + 1 and int(b[255]) == r + 1 for r, b in enumerate(blobs))
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.array([int(same), int(c.item()), len(pos), int(mine.sum()), int(blob_ok)] + bounds.tolist(), np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("res", [128, 256])
def test_two_ranks_agree_on_balanced_slabs(tmp_path, res):
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from flipsolver2d_b200 import scenes

    world = 2
    scene_path = scenes.write_scene(scenes.dam_break(res, "flip"), str(tmp_path / "scene.json"))
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, scene_path, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(str(tmp_path / ("r%d.npy" % r))) for r in range(world)]
    for g in got:
        same, total, seeds, mine, blob_ok = g[:5]
        assert same == 1 and blob_ok == 1
        assert total == seeds and seeds > 0
    bounds = got[0][5:]
    assert np.array_equal(bounds, got[1][5:])
    assert bounds[0] == 0 and bounds[-1] == res and all(b % 16 == 0 for b in bounds[1:-1])
    assert all(bounds[k + 1] - bounds[k] >= 32 for k in range(world))
    # balanced by particles, not by rows: the dam-break block sits in the lower half of the tank, so the cut lies below
    # the middle row and both ranks own a comparable number of particles
    assert bounds[1] > res // 2
    shares = [int(g[3]) for g in got]
    assert min(shares) > 0.25 * sum(shares), shares


@pytest.mark.parametrize("sim", ["flip", "smoke"])
def test_slab_seeding_keeps_the_own_rows_only(tmp_path, sim):
    """FlipSolver::seedRows: with a slab set, seedInitialFluid (and the smoke override) draws the WHOLE mt19937 stream but
    keeps only the particles of the rank's rows -- the shares of the ranks are disjoint, complete and identical to what
    the single solver seeds. Host only (no device is created before the first device call)."""
    sys.path.insert(0, ROOT)
    from flipsolver2d_b200 import host_api, scenes

    if sim == "flip":
        scene = scenes.dam_break(128, "flip")
    else:
        scene = scenes.smoke_test(128, parameter_handling="particle", sim_type="smoke")
        scene["solver"]["objects"].append({"type": "fluid", "viscosity": 0, "enabled": True, "verts": [[10, 10], [10, 40], [40, 40], [40, 10]]})
    path = scenes.write_scene(scene, str(tmp_path / ("seed_%s.json" % sim)))
    L = host_api.lib()
    K = 2 if sim == "flip" else 3

    def seeds(rank, world):
        L.fs2dh_set_quiet(1)
        L.fs2dh_set_slab(rank, world, 1)
        try:
            h = L.fs2dh_load_scene(path.encode())
            assert h
            assert L.fs2dh_prepare_host(h) == 0
            n = int(L.fs2dh_seed_count(h))
            pos = np.zeros((n, 2), np.float32)
            vel = np.zeros((n, 2), np.float32)
            props = np.zeros((K, n), np.float32)
            assert L.fs2dh_seed_particles(h, pos.ctypes.data_as(host_api.C.c_void_p), vel.ctypes.data_as(host_api.C.c_void_p),
                                          props.ctypes.data_as(host_api.C.c_void_p)) == 0
            bounds = np.zeros(world + 1, np.int32)
            assert L.fs2dh_slab_bounds(h, world, bounds.ctypes.data_as(host_api.C.c_void_p)) == 0
            L.fs2dh_destroy(h)
            return pos, bounds
        finally:
            L.fs2dh_set_slab(0, 1, 1)

    whole, _ = seeds(0, 1)
    assert len(whole) > 0
    world = 2
    parts = [seeds(r, world) for r in range(world)]
    bounds = parts[0][1]
    assert np.array_equal(bounds, parts[1][1])
    for r, (pos, _) in enumerate(parts):
        rows = np.floor(pos[:, 0]).astype(np.int64)
        assert np.all((rows >= bounds[r]) & (rows < bounds[r + 1]))
    joined = np.concatenate([p for p, _ in parts])
    assert len(joined) == len(whole)
    # same particles, same jitter: the single solver's seeds in row-major order are rank 0's followed by rank 1's only up to
    # the interleaving of rows, so compare as sets of bit patterns
    key = lambda a: np.sort(a.view(np.uint64).ravel())
    assert np.array_equal(key(np.ascontiguousarray(joined)), key(np.ascontiguousarray(whole)))
