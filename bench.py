#!/usr/bin/env python3
"""Headline benchmark: substeps/s of the FLIP dam-break scene at 4096^2 (BASELINE.json `metric`), one
JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res R] [--impl reference]

A "step" is one CFL substep = one FlipSolver::step() (flipsolver2d.cpp:412-462), driven through the
host mirror of the reference API (libfs2d_host.so: JsonSceneReader::loadJson -> stepSubstep), which
calls the sm_100a kernels through the fs2d C ABI. All state is resident in HBM when the timed region
starts (`value`); `e2e` repeats the measurement with the particle state crossing PCIe both ways every
step (pinned host buffers -> fs2d_upload_particles, substep, fs2d_download_particles).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libfs2d_ref.so = the
unmodified reference sources with its Release flags) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "substeps_per_s"
UNIT = "substeps/s"
REFERENCE_SAMPLE_RES = 1024


DENSE_FILL = False  # --fill dense: the near-full tank of BASELINE config 5 (~400 M particles at 8192^2)


def scene_for(res):
    from flipsolver2d_b200 import scenes
    return scenes.dam_break(res, "flip", ppc=8, pic_ratio=0.03, seed=0, max_substeps=10, dense_fill=DENSE_FILL)


def workload_name(res):
    return "flip dam-break %dx%d%s, ppc 8, picRatio 0.03, pcgIterLimit 200 (default), seed 0" % (
        res, res, " (near-full tank: fluid block (5,3)-(47,47) of the 50x50 domain)" if DENSE_FILL else "")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max((float(r[2]) for r in self.rows if len(r) > 2), default=None),
                "samples": len(self.rows), "reasons": sorted(reasons)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "pcg_traffic.json")))["solve_dram_bytes_per_launch"]
    except Exception:
        return None


# --------------------------------------------------------------------------------------- reference arm
def reference_substep(s):
    """One iteration of FlipSolver::stepFrame's loop on the oracle (flipsolver2d.cpp:476-497)."""
    st = reference_substep.state
    p = reference_substep.params
    vel = s.max_particle_velocity()
    import numpy as np
    f32 = np.float32
    max_dt = f32(p["cfl"]) / (f32(vel) + f32(1e-15))
    frame_dt = f32(p["frameDt"])
    finished = False
    if st["t"] + max_dt >= frame_dt or st["n"] == int(p["maxSubsteps"]) - 1:
        max_dt = frame_dt - st["t"]
        finished = True
    elif st["t"] + f32(2.0) * max_dt >= frame_dt:
        max_dt = f32(0.5) * (frame_dt - st["t"])
    s.set_step_dt(float(max_dt))
    s.stage("FULL_STEP")
    st["t"] = f32(st["t"] + max_dt)
    st["n"] += 1
    if finished:
        st["t"], st["n"] = f32(0.0), 0
        s.bump_frame()


def make_reference(res, tmp):
    import numpy as np
    from flipsolver2d_b200 import scenes
    from oracle import ref
    if not ref.available(strict=False):
        raise RuntimeError("oracle/_ref/libfs2d_ref.so missing (run __graft_entry__.build() where /root/reference exists)")
    path = scenes.write_scene(scene_for(res), os.path.join(tmp, "ref_%d.json" % res))
    s = ref.RefSolver(path, strict=False, threads=None, quiet=True)
    s.stage("FIRST_FRAME_INIT")
    s.bump_frame()
    reference_substep.state = {"t": np.float32(0.0), "n": 0}
    reference_substep.params = s.params()
    return s


def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref/libfs2d_ref.so: unmodified sources, Release flags, ThreadPool =
    all host threads) on the SAME scene at the SAME resolution as our arm. A 4096^2 substep costs the reference 30 - 200 s
    (SURVEY section 6), so `--steps K` of them do not fit a bench run: real substeps are run from frame 0 -- the first
    from rest, then moving ones -- until K are done or the wall budget (FS2D_REF_BUDGET_S, default 100 s: a new substep
    is only started while less than that has been spent) is used up, and the line reports the steps actually timed.
    Nothing is extrapolated; the 1024^2 sample the previous round scaled by 1/16 is kept as a secondary field."""
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="fs2d_bench_")
    budget = float(os.environ.get("FS2D_REF_BUDGET_S", "100"))
    t_init = time.perf_counter()
    s = make_reference(args.res, tmp)
    init_s = time.perf_counter() - t_init
    cores = s.threads
    particles = s.particle_count()
    durations = []
    t0 = time.perf_counter()
    while len(durations) < max(args.steps, 1) and (not durations or time.perf_counter() - t0 < budget):
        t1 = time.perf_counter()
        reference_substep(s)
        durations.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    steps = len(durations)
    stats = s.stats()
    s.close()
    value = steps / dt
    secondary = None
    if args.res != REFERENCE_SAMPLE_RES and os.environ.get("FS2D_REF_SECONDARY", "1") != "0":
        s2 = make_reference(REFERENCE_SAMPLE_RES, tmp)
        reference_substep(s2)
        t1 = time.perf_counter()
        n2 = 0
        while n2 < 2 or (time.perf_counter() - t1 < 10.0 and n2 < 8):
            reference_substep(s2)
            n2 += 1
        d2 = time.perf_counter() - t1
        s2.close()
        secondary = {"resolution": REFERENCE_SAMPLE_RES, "substeps": n2, "substeps_per_s": n2 / d2,
                     "note": "same scene at %dx%d, not scaled, not used for `value`" % (REFERENCE_SAMPLE_RES, REFERENCE_SAMPLE_RES)}
    sample = ("%d real substeps of the workload itself (%dx%d, %d particles) from frame 0 on %d host threads: %s s each "
              "(first from rest; later ones pay the reference's rebinParticles, which is superlinear in moved particles); "
              "scene set-up %.0f s not timed; requested --steps %d --warmup %d, stopped by the %.0f s wall budget"
              % (steps, args.res, args.res, particles, cores, "/".join("%.1f" % d for d in durations), init_s, args.steps,
                 args.warmup, budget))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 0, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.res), "cells": args.res * args.res, "particles": particles,
                       "pcg_iterations_last_frame": {"pressure": stats["pressure_iters"], "density": stats["density_iters"]}},
            "requested": {"steps": args.steps, "warmup": args.warmup},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "secondary_sample": secondary,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, tmp, budget_s=25.0):
    """The reference's CPU path on this host, bounded (~25 s): substeps of the same scene at 1024^2, rate scaled by the
    cell ratio. (`--impl reference` runs the workload itself at full size; that takes minutes.)"""
    res = REFERENCE_SAMPLE_RES
    s = make_reference(res, tmp)
    reference_substep(s)  # warm-up (first substep also pays first-touch allocation)
    t0 = time.perf_counter()
    n = 0
    while n < 2 or (time.perf_counter() - t0 < budget_s and n < 12):
        reference_substep(s)
        n += 1
    dt = time.perf_counter() - t0
    scale = (res * res) / float(args.res * args.res)
    out = {"value": n / dt * scale, "unit": UNIT, "cores": s.threads, "kind": "reference",
           "sample": "%d substeps of the same scene at %dx%d (%.1f s, %.0f ms/substep), rate x %.4f (cell ratio) to express it "
                     "at %dx%d -- favourable to the reference, whose rebin stage is superlinear; unmodified reference sources, "
                     "flags -O3 -mavx2 -ffast-math, ThreadPool = all host threads. The full-size measurement is "
                     "`bench.py --impl reference`."
                     % (n, res, res, dt, dt / n * 1e3, scale, args.res, args.res)}
    s.close()
    return out


# --------------------------------------------------------------------------------------- multi-GPU self checks
MG_PARITY_RES = 1024
MG_PARITY_STEPS = 6


def mg_parity_check(make_solver, total, world, rank, torch, dist):
    """The multi-process slab path against a single handle, inside the bench run: every rank steps (a) its slab of a
    1024^2 dam break together with the other ranks and (b) a private single-handle solver of the same scene on its own
    GPU, both for MG_PARITY_STEPS substeps from rest with BASELINE's settings (density 0.5: both PCG solves of every
    substep run into the 200-iteration cap, i.e. a fixed iteration count), then compares the rows it owns: particle
    count, material grid (bit-exact), U and pressure (relative L2: the dot-product partials are grouped by rank, which
    is the only arithmetic the decomposition changes)."""
    import numpy as np
    scene = scene_for(MG_PARITY_RES)
    slab = make_solver(scene, "mgp_slab")
    single = make_solver(scene, "mgp_single", slab=False)
    slab.prepare()
    single.prepare()
    for _ in range(MG_PARITY_STEPS):
        slab.step_substep()
        single.step_substep()
    ds, d1 = slab.device(num_properties=2), single.device(num_properties=2)
    lo, hi, _ = ds.slab_rows()
    I, J = slab.I, slab.J
    p1, _, _ = d1.download_particles()
    own1 = int(((np.floor(p1[:, 0]) >= lo) & (np.floor(p1[:, 0]) < hi)).sum())

    def rows(a, per_row):
        return a.reshape(-1, per_row)[lo:hi]

    mat_equal = bool(np.array_equal(rows(ds.download("MATERIAL"), J), rows(d1.download("MATERIAL"), J)))
    cnt_equal = bool(np.array_equal(rows(ds.download("COUNTS"), J), rows(d1.download("COUNTS"), J)))

    def sq(name, per_row):
        a, b = rows(ds.download(name), per_row).astype(np.float64), rows(d1.download(name), per_row).astype(np.float64)
        return float(((a - b) ** 2).sum()), float((b ** 2).sum()), float(a.sum()), float(b.sum())

    u, p = sq("U", J), sq("PRESSURE", J)
    own_slab = slab.particle_count()
    out = {"resolution": MG_PARITY_RES, "substeps": MG_PARITY_STEPS, "ranks": world,
           "particles_slab_total": total(own_slab), "particles_single": single.particle_count(),
           "particles_owned_rows_equal": bool(total(int(own_slab == own1)) == world),
           "material_rows_equal": bool(total(int(mat_equal)) == world), "counts_rows_equal": bool(total(int(cnt_equal)) == world),
           "u_rel_l2": (total(u[0]) / max(total(u[1]), 1e-300)) ** 0.5, "pressure_rel_l2": (total(p[0]) / max(total(p[1]), 1e-300)) ** 0.5,
           "sum_u": {"slab": total(u[2]), "single": total(u[3])}, "sum_pressure": {"slab": total(p[2]), "single": total(p[3])},
           "pcg_iterations": {"slab": slab.stats()["pressure_iters"], "single": single.stats()["pressure_iters"]}}
    out["ok"] = bool(out["particles_slab_total"] == out["particles_single"] and out["particles_owned_rows_equal"]
                     and out["material_rows_equal"] and out["counts_rows_equal"] and out["u_rel_l2"] < 1e-5
                     and out["pressure_rel_l2"] < 1e-5)
    slab.close()
    single.close()
    return out


def config5_region(make_solver, timed, total, world, local_rank, torch, capi, args):
    """BASELINE config 5 inside the N-GPU run: the flip dam break at 8192^2 with the near-full tank (fluid block
    (5,3)-(47,47): ~397 M particles), slab-decomposed over the N GPUs; a few substeps after warm-up, timed like `value`."""
    global DENSE_FILL
    keep = DENSE_FILL
    DENSE_FILL = args.config5_fill == "dense"
    try:
        t0 = time.perf_counter()
        sv = make_solver(scene_for(8192), "config5")
        sv.prepare()
        setup_s = time.perf_counter() - t0
        d = sv.device(num_properties=2)
        st = torch.cuda.ExternalStream(capi.lib().fs2d_stream(d.h), device=torch.device("cuda", local_rank))
        for _ in range(2):
            sv.step_substep()
        steps = args.config5_steps
        ms = timed(sv.step_substep, steps, d=d, on=st)
        stats = sv.stats()
        out = {"workload": workload_name(8192), "cells": 8192 * 8192, "particles": total(sv.particle_count()), "n_gpus": world,
               "steps": steps, "warmup": 2, "value": steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
               "pcg_active_cells": total(int(d.pcg_active_cells())), "setup_s": round(setup_s, 1),
               "pcg_iterations_last_frame": {"pressure": stats["pressure_iters"], "density": stats["density_iters"]}}
        sv.close()
        return out
    finally:
        DENSE_FILL = keep


# --------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from flipsolver2d_b200 import capi, host_api, scenes

    if capi.lib().fs2d_device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    tmp = tempfile.mkdtemp(prefix="fs2d_bench_")
    res = args.res

    def make_solver(scene, tag, slab=True):
        """JsonSceneReader::loadJson on every rank. slab: ONE scene decomposed into row slabs, one rank per GPU (strong
        scaling): the ranks exchange halo rows, migrating particles and the PCG reduction partials through peer-mapped
        device memory (CUDA IPC over NVLink); torch.distributed only carries the 256-byte IPC blobs at start-up and the
        timing reductions."""
        path = scenes.write_scene(scene, os.path.join(tmp, "%s_r%d.json" % (tag, rank)))
        sv = host_api.Solver(path, quiet=True, device=local_rank, slab=(rank, world) if (world > 1 and slab) else None)
        if world > 1 and slab:
            blob = torch.frombuffer(bytearray(sv.slab_export()), dtype=torch.uint8).cuda()
            blobs = [torch.empty_like(blob) for _ in range(world)]
            dist.all_gather(blobs, blob)
            for r in range(world):
                if r != rank:
                    sv.slab_connect(r, bytes(blobs[r].cpu().numpy().tobytes()))
            dist.barrier()
        return sv

    solver = make_solver(scene_for(res), "scene_%d" % res)
    solver.prepare()  # frame-0 rasterisation + seeding + upload: set-up, not timed
    dev = solver.device(num_properties=2)
    N = solver.N
    rows_lo, rows_hi = (0, solver.I)
    if world > 1:
        rows_lo, rows_hi, _ = dev.slab_rows()
    own_cells = (rows_hi - rows_lo) * solver.J

    def total(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return type(v)(t.item())
    stream = torch.cuda.ExternalStream(capi.lib().fs2d_stream(dev.h), device=torch.device("cuda", local_rank))

    def barrier(d=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        (d or dev).synchronize()

    def timed(fn, steps, d=None, on=None):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the solver's stream, max over ranks."""
        barrier(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(on or stream)
        for _ in range(steps):
            fn()
        e1.record(on or stream)
        barrier(d)
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- resident run
    for _ in range(max(args.warmup, 3)):
        solver.step_substep()
    dev.pcg_profile(True)
    dev.kernel_profile(True)
    launches0 = solver.kernel_launches()
    particles_before = total(solver.particle_count())
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(solver.step_substep, args.steps)
    launches = total(solver.kernel_launches() - launches0)
    prof_ms, prof_n = dev.pcg_profile_read()
    solve_ms, solve_n = dev.pcg_profile_solves()
    dev.pcg_profile(False)
    kgroups = dev.kernel_profile_read()
    dev.kernel_profile(False)
    stats = solver.stats()
    particles = solver.particle_count()          # this rank's
    particles_all = total(particles)
    active_cells = dev.pcg_active_cells()        # this rank's
    active_cells_all = total(int(active_cells))

    # ---- the same substeps with the PCG kernels walking the WHOLE grid (fs2d_pcg_set_dense): every vector
    # pass streams 128 MB from HBM, which is the configuration the HBM roofline of SURVEY 8(d) is defined on
    dev.pcg_set_dense(True)
    solver.step_substep()
    dev.pcg_profile(True)
    dense_ms = timed(solver.step_substep, args.steps)
    dprof_ms, dprof_n = dev.pcg_profile_read()
    dsolve_ms, dsolve_n = dev.pcg_profile_solves()
    clocks = sampler.summary()  # sampled across both timed regions (active-tile walk and dense walk)
    dev.pcg_profile(False)
    dev.pcg_set_dense(False)
    solver.step_substep()

    # ---- end to end: the particle state crosses PCIe both ways every step, through the packed C-ABI calls (one copy per
    # direction; lossless: the storage-bin byte travels with the records, tests/test_substep_gpu.py proves the round trip
    # is the identity on the solver state)
    # The region runs on a SECOND solver of the same scene, warmed up by the same number of substeps, so that it covers
    # the same simulation window as `value` (the dam break spreads: later substeps have more active PCG tiles and a
    # larger extrapolation box; timing the regions back to back on one solver compared different work).
    if not args.no_e2e:
        resident_solver, resident_dev = solver, dev
        solver = make_solver(scene_for(res), "scene_e2e_%d" % res)
        solver.prepare()
        dev = solver.device(num_properties=2)
        stream = torch.cuda.ExternalStream(capi.lib().fs2d_stream(dev.h), device=torch.device("cuda", local_rank))
        for _ in range(max(args.warmup, 3) - 1):   # + the warm-up step of the transfer path below
            solver.step_substep()
    P = solver.particle_count()
    K = 2
    rec = 16 + 4 * K + 1
    cap = int(P * 1.25) + 1024
    if world > 1:
        cap = int(P * 1.5) + 2_000_000  # slabs: particles migrate between ranks (a few rows per step at CFL 5)
    if args.no_e2e:
        cap = 16
    hostbuf = torch.empty((cap * rec,), dtype=torch.uint8).pin_memory()
    L = capi.lib()
    import ctypes
    state = {"n": P, "h2d": 0, "d2h": 0, "n_min": P, "n_max": P}

    def download():
        n = ctypes.c_int64(0)
        rc = L.fs2d_download_particles_packed(dev.h, hostbuf.data_ptr(), hostbuf.numel(), ctypes.byref(n))
        assert rc == 0, dev.L.fs2d_last_error(dev.h)
        state["n"] = n.value
        state["n_min"] = min(state["n_min"], n.value)
        state["n_max"] = max(state["n_max"], n.value)
        state["d2h"] += n.value * rec

    phase_s = {"upload": 0.0, "substep": 0.0, "download": 0.0}
    copy_ms = {}

    # one handle: the streamed calls (sections of the pinned buffer travel on a copy stream while the substep runs:
    # FlipSolver::stepSubstepStreamed); row slabs: packed upload -> substep -> packed download
    streamed = world == 1

    def download_sectioned():
        n = dev.stream_end(hostbuf.numpy(), cap)
        state["n"] = n
        state["n_min"] = min(state["n_min"], n)
        state["n_max"] = max(state["n_max"], n)

    def e2e_step_streamed():
        n = state["n"]
        t0 = time.perf_counter()
        _, m = solver.step_substep_streamed(hostbuf.data_ptr(), cap, n)
        t1 = time.perf_counter()
        for k, v in dev.stream_timing().items():   # CUDA-event durations of this step's copies (read after the step)
            copy_ms[k] = copy_ms.get(k, 0.0) + v
        state["h2d"] += n * rec
        state["d2h"] += m * rec
        state["n"] = m
        state["n_min"] = min(state["n_min"], m)
        state["n_max"] = max(state["n_max"], m)
        phase_s["substep"] += t1 - t0

    def e2e_step():
        if streamed:
            return e2e_step_streamed()
        n = state["n"]
        t0 = time.perf_counter()
        rc = L.fs2d_upload_particles_packed(dev.h, hostbuf.data_ptr(), n)
        assert rc == 0, dev.L.fs2d_last_error(dev.h)
        state["h2d"] += n * rec
        t1 = time.perf_counter()
        solver.step_substep()
        dev.synchronize()
        t2 = time.perf_counter()
        download()
        t3 = time.perf_counter()
        phase_s["upload"] += t1 - t0
        phase_s["substep"] += t2 - t1
        phase_s["download"] += t3 - t2

    e2e = None
    if not args.no_e2e:
        if streamed:
            download_sectioned()
        else:
            download()
        e2e_step()  # warm-up of the path
        state["h2d"] = state["d2h"] = 0
        for k in phase_s:
            phase_s[k] = 0.0
        copy_ms.clear()
        state["n_min"] = state["n_max"] = n_start = state["n"]
        e2e_steps = args.steps
        dev.pcg_profile(True)
        dev.kernel_profile(True)
        e2e_ms = timed(e2e_step, e2e_steps)
        e2e_solve_ms, e2e_solve_n = dev.pcg_profile_solves()
        e2e_groups = dev.kernel_profile_read()
        dev.pcg_profile(False)
        dev.kernel_profile(False)
        e2e = {"value": e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": total(state["h2d"]) // e2e_steps,
               "d2h_bytes_per_step": total(state["d2h"]) // e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
               "particles_start": total(n_start), "particles_end": total(state["n"]), "bytes_per_particle": rec,
               "host_ms_per_step_rank0": {k: round(v / e2e_steps * 1e3, 3) for k, v in phase_s.items()},
               "copy_ms_per_step_rank0": {k: round(v / e2e_steps, 3) for k, v in copy_ms.items()},
               "pcg_solve_kernel_ms": e2e_solve_ms / max(e2e_solve_n, 1),
               "kernel_group_ms": {g: round(v[0] / max(v[1], 1), 3) for g, v in e2e_groups.items()},
               "stage_ms_per_substep_last_frame": (lambda st: {n: round(float(st["timings"][k]) / max(st["substeps"], 1), 3)
                                                               for k, n in enumerate(host_api.STAGES)})(solver.stats()),
               "window": "substeps %d .. %d of the scene, the window `value` was timed on (second solver, same warm-up)" % (max(args.warmup, 3) + 1, max(args.warmup, 3) + 1 + e2e_steps),
               "what": ("per step: FlipSolver::stepSubstepStreamed on a pinned host buffer = fs2d_particle_stream_begin (all %d "
                        "bytes per record host -> device on a copy stream, each section awaited where a stage first reads it), the "
                        "stages of stepSubstep, fs2d_particle_stream_positions_final after the density correction (positions and "
                        "property columns device -> host under the P2G / pressure stages), fs2d_particle_stream_end (velocities, "
                        "storage bytes, reseeded records; returns when the buffer is complete)" % rec) if streamed else
                       "per step: fs2d_upload_particles_packed from a pinned host buffer, FlipSolver::stepSubstep, "
                       "fs2d_download_particles_packed into it (positions, velocities, %d property columns, storage-bin byte)" % K
                       + (" (every rank moves the particles of its slab; bytes summed over ranks)" if world > 1 else "")}

    # ---- N > 1: the run proves itself and carries BASELINE config 5
    mg_parity = mg_parity_check(make_solver, total, world, rank, torch, dist) if world > 1 and not args.no_mg_parity else None
    config_5 = None
    if world > 1 and not args.no_config5:
        solver.close()
        if not args.no_e2e:
            resident_solver.close()
        del hostbuf
        torch.cuda.empty_cache()
        config_5 = config5_region(make_solver, timed, total, world, local_rank, torch, capi, args)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel, pcgSolveKernel: ONE cooperative launch per PCG solve runs all iterations
    # (two solves per substep: density correction + pressure projection). Per iteration and cell it moves, in its K1
    # phase (s = z + beta s, x += alpha s, q = A s, q.s) R z,s,x + W s,q,x + 1 B row info = 49 B and in its K2 phase
    # (r -= alpha q, z = M r, z.r, max|r|) R r,q + W r,z + 2 B preconditioner info = 34 B: 83 B.
    # Launch durations come from CUDA events around every launch on the solver's stream (fs2d_pcg_profile_solves);
    # the K1 / K2 split of a launch from the kernel's own phase clocks (%globaltimer in CTA 0).
    peak, peak_src = measured_peak()

    def solve_roofline(p_ms, p_n, s_ms, s_n, cells):
        launches = max(int(s_n), 1)
        iters = int(p_n[0]) / launches
        launch_ms = s_ms / launches
        bytes_per_launch = 83.0 * cells * iters
        k1_ms = p_ms[0] / max(int(p_n[0]), 1)
        k2_ms = p_ms[1] / max(int(p_n[1]), 1)
        return {"cells_per_launch": int(cells), "iterations_per_launch": iters, "bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": launch_ms, "launches_timed": int(s_n),
                "achieved": bytes_per_launch / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0,
                "k1_phase": {"bytes_per_iteration": 49 * cells, "avg_ms": k1_ms,
                             "achieved": 49 * cells / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0},
                "k2_phase": {"bytes_per_iteration": 34 * cells, "avg_ms": k2_ms,
                             "achieved": 34 * cells / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0}}

    dense = solve_roofline(dprof_ms, dprof_n, dsolve_ms, dsolve_n, own_cells)
    act = solve_roofline(prof_ms, prof_n, solve_ms, solve_n, active_cells)
    for r in (dense, act):
        for k in ("k1_phase", "k2_phase"):
            r[k]["frac"] = r[k]["achieved"] / peak
    roofline = {"bound": "hbm", "kernel": "pcgSolveKernel", "achieved": dense["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dense["achieved"] / peak, "traffic": ncu_traffic() if (world == 1 and res == 4096 and not DENSE_FILL) else None,
                "traffic_source": "profiles/pcg_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                                  "capture of this kernel on this workload (1 GPU, dense walk); null on any other configuration",
                "peak_source": peak_src,
                "bytes_per_launch": dense["bytes_per_launch"], "avg_launch_ms": dense["avg_launch_ms"],
                "launches_timed": dense["launches_timed"], "iterations_per_launch": dense["iterations_per_launch"],
                "k1_phase": dense["k1_phase"], "k2_phase": dense["k2_phase"],
                "measured_in": "second timed region of this run: the same %d substeps with fs2d_pcg_set_dense(1), i.e. the "
                               "kernel walks all %d cells%s and every vector pass comes from HBM" %
                               (args.steps, own_cells, " of rank 0's slab (the launch includes the cross-GPU barriers that "
                                "carry the all-reduces and the halo rows)" if world > 1 else ""),
                "pcg_share_of_step": dsolve_ms / dense_ms if dense_ms > 0 else None,
                "dense_walk": {"value": args.steps / (dense_ms * 1e-3), "unit": UNIT, "ms_per_step": dense_ms / args.steps},
                "active_tile_walk": dict(act, pcg_share_of_step=solve_ms / ms if ms > 0 else None,
                                         note="default mode (timed region of `value`): tiles without matrix rows are skipped; "
                                              "the %d walked cells x 7 vectors fit the 126 MB L2, so these GB/s are not an HBM "
                                              "figure" % active_cells)}
    # ---- transfer kernels: algorithmic bytes of SURVEY 8(d) (P particles, N cells, k float property columns; one centred
    # parameter, the viscosity) over the device time of each kernel group in the timed region of `value` (CUDA events on
    # the solver's stream around every call, fs2d_kernel_profile). These kernels are gathers with the reference's exact
    # float expressions: ncu (profiles/r2_transfer_ncu_full_summary.csv) shows them bound by instruction issue / latency
    # at low occupancy, not by DRAM -- the fractions say how far from the HBM roofline that leaves them.
    Pn, Nn, kprops = float(particles), float(own_cells), 2.0
    transfer_bytes = {"SORT": 2 * (16 + 4 * kprops) * Pn + 12 * Pn + 12 * Nn,
                      "P2G": (16 + 4 * kprops) * Pn + (10 + 5 * 1) * Nn + 4 * Nn,
                      "SDF": 8 * Pn + 4 * Nn + 4 * Nn + 2 * Nn,
                      "DENSITY": 8 * Pn + 4 * Nn + 4 * Nn,
                      "ADVECT+G2P": 32 * Pn + 20 * Nn}
    kg = dict(kgroups)
    kg["ADVECT+G2P"] = (kg["ADVECT"][0] + kg["G2P"][0], kg["ADVECT"][1])
    roofline_transfer = {"peak": peak, "unit": "GB/s", "particles": int(Pn), "cells": int(Nn),
                         "formulas": "SURVEY.md 8(d): sort 2(16+4k)P+12P+12N; P2G (16+4k)P+(10+5)N+4N; sdf 8P+10N; density 8P+8N; "
                                     "G2P + RK4 advect 32P+20N (k = 2 property columns)", "groups": {}}
    for name, nbytes in transfer_bytes.items():
        t_ms, calls = kg[name]
        if calls > 0 and t_ms > 0:
            per_call = t_ms / calls
            roofline_transfer["groups"][name] = {"bytes_per_call": nbytes, "avg_ms": per_call, "calls": calls,
                                                 "achieved": nbytes / (per_call * 1e-3) / 1e9,
                                                 "frac": nbytes / (per_call * 1e-3) / 1e9 / peak}
    roofline_transfer["groups"]["EXTRAPOLATE (velocity BFS, 2 calls per substep)"] = {
        "avg_ms": kg["EXTRAPOLATE"][0] / max(kg["EXTRAPOLATE"][1], 1), "calls": kg["EXTRAPOLATE"][1]}
    per = max(stats["substeps"], 1)
    stage_ms = {name: round(float(stats["timings"][k]) / per, 3) for k, name in enumerate(host_api.STAGES)}
    base = cpu_baseline(args, tmp) if world == 1 and not args.no_cpu_baseline else None
    line = {"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong",  # one scene of fixed size on 1 .. N GPUs
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(res), "cells": N, "particles": particles_all,
                       "particles_timed_region": {"start": particles_before, "end": particles_all},
                       "l2": "every PCG vector (%d MB%s) and the particle arrays exceed the 126 MB L2"
                             % (own_cells * 8 // 2 ** 20, " per rank" if world > 1 else ""),
                       "parallelism": "1 GPU" if world == 1 else
                       "%d row slabs of one scene, one rank per GPU; halo rows, particle migration and the PCG all-reduces "
                       "go through peer-mapped memory (CUDA IPC over NVLink); the all-reduces are the grid barriers of the whole-solve "
                       "kernel; slab boundaries balanced by particles per tile row" % world,
                       "pcg_iterations_last_frame": {"pressure": stats["pressure_iters"], "density": stats["density_iters"]},
                       "pcg_walk": "active tiles (%d of %d cells)" % (active_cells_all, N),
                       "stage_ms_per_substep_last_frame": stage_ms},
            "roofline": roofline, "roofline_transfer": roofline_transfer, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks}
    if mg_parity is not None:
        line["mg_parity"] = mg_parity
    if config_5 is not None:
        line["config_5"] = config_5
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fill", default="reference", choices=["reference", "dense"])
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end region (very large particle counts)")
    ap.add_argument("--no-mg-parity", action="store_true", help="N > 1: skip the slab-vs-single-handle self check")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the 8192^2 region (BASELINE config 5)")
    ap.add_argument("--config5-fill", default="dense", choices=["reference", "dense"])
    ap.add_argument("--config5-steps", type=int, default=3)
    args = ap.parse_args()
    global DENSE_FILL
    DENSE_FILL = args.fill == "dense"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
