#!/usr/bin/env python3
"""BASELINE config 4: nbflip with viscosityEnabled at 4096^2, slab-decomposed over the GPUs of one box.

  python tools/run_config4.py [--res 4096] [--steps 10]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N tools/run_config4.py    (N GPUs, one rank per GPU)

One JSON line (rank 0): substeps/s (CUDA events on the solver stream, max over ranks), stage times, iteration counts,
and a parity block: every rank also steps a private single-handle solver of the same scene at --check-res and compares
its rows (particle count, material grid, U, pressure)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from flipsolver2d_b200 import capi, host_api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check-res", type=int, default=512)
    ap.add_argument("--check-steps", type=int, default=6)
    ap.add_argument("--no-viscosity", action="store_true")
    ap.add_argument("--scene", default="nbflip", choices=["nbflip", "smoke", "fire"],
                    help="smoke / fire: BASELINE config 3's scene (grid-advected parameters) over the same row slabs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    capi.lib()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix="fs2d_c4_")

    def total(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return type(v)(t.item())

    def make(res, tag, slab=True):
        if args.scene == "nbflip":
            sc = scenes.dam_break(res, "nbflip", viscosity_enabled=not args.no_viscosity)
        else:
            sc = scenes.smoke_test(res, ppc=4, parameter_handling="grid", sim_type=args.scene)
        path = scenes.write_scene(sc, os.path.join(tmp, "%s_r%d.json" % (tag, rank)))
        sv = host_api.Solver(path, quiet=True, device=local, slab=(rank, world) if (world > 1 and slab) else None)
        if world > 1 and slab:
            blob = torch.frombuffer(bytearray(sv.slab_export()), dtype=torch.uint8).cuda()
            blobs = [torch.empty_like(blob) for _ in range(world)]
            dist.all_gather(blobs, blob)
            for r in range(world):
                if r != rank:
                    sv.slab_connect(r, bytes(blobs[r].cpu().numpy().tobytes()))
            dist.barrier()
        return sv

    parity = None
    if world > 1:
        slab, single = make(args.check_res, "chk_slab"), make(args.check_res, "chk_single", slab=False)
        slab.prepare()
        single.prepare()
        for _ in range(args.check_steps):
            slab.step_substep()
            single.step_substep()
        ds, d1 = slab.device(2), single.device(2)
        lo, hi, _ = ds.slab_rows()
        J = slab.J

        def rows(a, per):
            return a.reshape(-1, per)[lo:hi]

        def rel(name, per):
            a, b = rows(ds.download(name), per).astype(np.float64), rows(d1.download(name), per).astype(np.float64)
            return float(((a - b) ** 2).sum()), float((b ** 2).sum())

        u, p = rel("U", J), rel("PRESSURE", J)
        mat = bool(np.array_equal(rows(ds.download("MATERIAL"), J), rows(d1.download("MATERIAL"), J)))
        parity = {"resolution": args.check_res, "substeps": args.check_steps, "particles_slab_total": total(slab.particle_count()),
                  "particles_single": single.particle_count(), "material_rows_equal": bool(total(int(mat)) == world),
                  "u_rel_l2": (total(u[0]) / max(total(u[1]), 1e-300)) ** 0.5,
                  "pressure_rel_l2": (total(p[0]) / max(total(p[1]), 1e-300)) ** 0.5,
                  "viscosity_iters": {"slab": slab.stats()["viscosity_iters"], "single": single.stats()["viscosity_iters"]}}
        slab.close()
        single.close()

    t0 = time.perf_counter()
    sv = make(args.res, "c4")
    sv.prepare()
    setup_s = time.perf_counter() - t0
    d = sv.device(2)
    stream = torch.cuda.ExternalStream(capi.lib().fs2d_stream(d.h), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        d.synchronize()

    for _ in range(args.warmup):
        sv.step_substep()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sv.step_substep()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    st = sv.stats()
    per = max(st["substeps"], 1)
    name = "c4_nbflip%d_%s" % (args.res, "inviscid" if args.no_viscosity else "viscous") if args.scene == "nbflip" else "c3_%s%d_grid" % (args.scene, args.res)
    out = {"config": name, "n_gpus": world,
           "cells": sv.N, "particles": total(sv.particle_count()), "substeps_per_s": args.steps / (ms * 1e-3),
           "ms_per_substep": ms / args.steps, "steps": args.steps, "warmup": args.warmup, "setup_s": round(setup_s, 1),
           "iterations_last_frame": {"pressure": st["pressure_iters"], "viscosity": st["viscosity_iters"]},
           "stage_ms_per_substep_last_frame_rank0": {n: round(float(st["timings"][k]) / per, 3) for k, n in enumerate(host_api.STAGES)},
           "pcg_active_cells": total(int(d.pcg_active_cells())), "parity": parity}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
