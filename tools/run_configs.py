#!/usr/bin/env python3
"""Run the BASELINE.json configurations that fit one GPU through the host mirror (JsonSceneReader -> stepSubstep)
and print one JSON line per configuration: substeps/s (CUDA events around the timed substeps), per-stage ms, PCG /
viscosity iteration counts, particles.

  python tools/run_configs.py [--steps K] [--warmup W] [--only name,...]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flipsolver2d_b200 import capi, host_api, scenes  # noqa: E402

CONFIGS = {
    # BASELINE.json `configs`, in order
    "c1_flip128": lambda: scenes.dam_break(128, "flip", ppc=8),
    "c2_flip1024": lambda: scenes.dam_break(1024, "flip", ppc=8, pic_ratio=0.03),
    "c3_smoke2048_grid": lambda: scenes.smoke_test(2048, ppc=4, parameter_handling="grid"),
    "c4_nbflip4096_viscous_1gpu": lambda: scenes.dam_break(4096, "nbflip", viscosity_enabled=True),
    "c5_flip8192_1gpu": lambda: scenes.dam_break(8192, "flip", ppc=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    capi.lib()
    torch.cuda.set_device(0)
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    tmp = tempfile.mkdtemp(prefix="fs2d_cfg_")
    for name in names:
        path = scenes.write_scene(CONFIGS[name](), os.path.join(tmp, name + ".json"))
        t0 = time.time()
        s = host_api.Solver(path, quiet=True)
        s.prepare()
        setup_s = time.time() - t0
        props = {0: 2, 1: 3, 2: 4, 3: 2}[capi.lib().fs2d_device_count() and s.L.fs2dh_sim_type(s.h)]
        dev = s.device(num_properties=props)
        stream = torch.cuda.ExternalStream(capi.lib().fs2d_stream(dev.h))
        for _ in range(args.warmup):
            s.step_substep()
        dev.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            s.step_substep()
        e1.record(stream)
        dev.synchronize()
        ms = e0.elapsed_time(e1)
        st = s.stats()
        per = max(st["substeps"], 1)
        print(json.dumps({"config": name, "cells": s.N, "particles": s.particle_count(), "substeps_per_s": args.steps / (ms * 1e-3),
                          "ms_per_substep": ms / args.steps, "steps": args.steps, "warmup": args.warmup, "setup_s": round(setup_s, 2),
                          "iterations_last_frame": {"pressure": st["pressure_iters"], "density": st["density_iters"],
                                                    "viscosity": st["viscosity_iters"]},
                          "stage_ms_per_substep_last_frame": {n: round(float(st["timings"][k]) / per, 3)
                                                              for k, n in enumerate(host_api.STAGES)},
                          "kernel_launches": s.kernel_launches()}), flush=True)
        s.close()


if __name__ == "__main__":
    main()
