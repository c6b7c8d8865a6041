#!/usr/bin/env python3
"""2-GPU timing probe of the slab PCG (run under torchrun, one rank per GPU): per-kernel average launch times of
200-iteration solves on the 4096^2 dam-break matrix, dense and active-tile walks, with the FS2D_MG_DEBUG switches.

  torchrun --nproc-per-node 2 tools/mg_probe.py [res] [solves]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from flipsolver2d_b200 import capi, host_api, scenes  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
res = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
capi.lib()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = scenes.dam_break(res, "flip")
path = scenes.write_scene(sc, "/tmp/mgprobe_%d.json" % rank)
hs = host_api.Solver(path, quiet=True, device=local)
hs.prepare_host()
mat = hs.host_grid("MATERIAL")
I, J = hs.I, hs.J
d = capi.Device(I, J, dx=50.0 / res, fluid_density=0.5, device=local)
if world > 1:
    # FS2D_PROBE_BALANCED=1: slab boundaries cut by particles per tile row (what bench.py / the host mirror use), so the
    # boundaries run through the fluid and halo rows are really pushed; default: equal row counts
    bounds = hs.slab_bounds(world) if os.environ.get("FS2D_PROBE_BALANCED") else None
    d.slab_configure(rank, world, row_bounds=bounds)
    out_bounds = None if bounds is None else [int(b) for b in bounds]
    blob = torch.frombuffer(bytearray(d.slab_export()), dtype=torch.uint8).cuda()
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    for r in range(world):
        if r != rank:
            d.slab_connect(r, blobs[r].cpu().numpy().tobytes())
    dist.barrier()
d.upload("MATERIAL", mat)
d.set_step_dt(1.0 / 300.0)
d.stage("build_matrix")
unit = (mat == capi.FLUID)
rng = np.random.default_rng(3)
rhs = np.where(unit, rng.standard_normal(I * J), 0.0)
d.upload("RHS", rhs)
out = {"rank": rank, "world": world, "res": res, "debug": os.environ.get("FS2D_MG_DEBUG", "0"),
       "resident": os.environ.get("FS2D_PCG_RESIDENT", "1"), "bounds": out_bounds if world > 1 else None}
for dense in (True, False):
    d.pcg_set_dense(dense)
    d.pcg_solve_device(200, 0.0)
    d.synchronize()
    d.pcg_profile(True)
    if world > 1:
        dist.barrier()
    for _ in range(solves):
        d.upload("RHS", rhs)
        d.pcg_solve_device(200, 0.0)
    d.synchronize()
    ms, n = d.pcg_profile_read()
    d.pcg_profile(False)
    key = "dense" if dense else "active"
    out[key] = {"k1_us": 1e3 * ms[0] / max(n[0], 1), "k2_us": 1e3 * ms[1] / max(n[1], 1), "launches": int(n[0]),
                "iters": d.pcg_last_iterations(), "active_cells": int(d.pcg_active_cells())}
import time  # noqa: E402
for dense in (True, False):
    d.pcg_set_dense(dense)
    d.pcg_solve_device(200, 0.0)
    d.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(solves):
        d.pcg_solve_device(200, 0.0)
    t1 = time.perf_counter()
    d.synchronize()
    t2 = time.perf_counter()
    out["noprof_" + ("dense" if dense else "active")] = {"host_enqueue_us_per_iter": 1e6 * (t1 - t0) / (200 * solves),
                                                         "total_us_per_iter": 1e6 * (t2 - t0) / (200 * solves)}
if int(os.environ.get("FS2D_MG_DEBUG", "0")) & 8:
    # whole-solve kernel timeline (CTA 0 + the last-arriving CTA, globaltimer ns): per phase
    # 0 walk start, 1 first tile landed, 2 walk end, 3 CTA 0 arrived, 4 last CTA arrived, 5 published, 6 CTA 0 released
    import ctypes as C
    L = capi.lib()
    L.fs2d_debug_mg_timeline.argtypes = [C.c_void_p, C.c_void_p]
    for dense in (True, False):
        d.pcg_set_dense(dense)
        d.pcg_solve_device(200, 0.0)
        d.synchronize()
        tl = np.zeros((1024, 8), np.uint64)
        L.fs2d_debug_mg_timeline(d.h, tl.ctypes.data_as(C.c_void_p))
        t = tl[101:301].astype(np.int64)  # phases 101..300: K1 odd, K2 even
        rows = {}
        if d.pcg_last_kernel() > 0:
            # pcgResidentKernel stamps: 0 phase start, 1 resident tiles done, 2 paged tiles done, 3 barrier passed (CTA 0)
            for name, sel in (("k1", t[0::2]), ("k2", t[1::2])):
                rows[name] = {"resident_us": float(np.mean(sel[:, 1] - sel[:, 0])) / 1e3, "paged_us": float(np.mean(sel[:, 2] - sel[:, 1])) / 1e3,
                              "barrier_us": float(np.mean(sel[:, 3] - sel[:, 2])) / 1e3,
                              "barrier_p10_p50_p90_us": [float(x) / 1e3 for x in np.percentile(sel[:, 3] - sel[:, 2], [10, 50, 90])]}
            rows["phase_period_us"] = float(np.mean(np.diff(t[:, 0]))) / 1e3
            rows["kernel"] = int(d.pcg_last_kernel())
            out["timeline_" + ("dense" if dense else "active")] = rows
            continue
        for name, sel in (("k1", t[0::2]), ("k2", t[1::2])):
            rows[name] = {"first_tile_us": float(np.mean(sel[:, 1] - sel[:, 0])) / 1e3,
                          "walk_rest_us": float(np.mean(sel[:, 2] - sel[:, 1])) / 1e3,
                          "cta0_reduce_arrive_us": float(np.mean(sel[:, 3] - sel[:, 2])) / 1e3,
                          "wait_for_last_cta_us": float(np.mean(sel[:, 4] - sel[:, 3])) / 1e3,
                          "final_reduce_publish_us": float(np.mean(sel[:, 5] - sel[:, 4])) / 1e3,
                          "release_us": float(np.mean(sel[:, 6] - sel[:, 5])) / 1e3,
                          "phase_us": float(np.mean(sel[:, 6] - sel[:, 0])) / 1e3}
        rows["phase_period_us"] = float(np.mean(np.diff(t[:, 0]))) / 1e3
        out["timeline_" + ("dense" if dense else "active")] = rows
print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
