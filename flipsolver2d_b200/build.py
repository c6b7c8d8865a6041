#!/usr/bin/env python3
"""Build libfs2d_cuda.so (all of csrc/*.cu) for sm_100a with nvcc, in-tree.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the gpurun snapshot. cudart is linked statically (nvcc default) so the
library only needs the driver at run time.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfs2d_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr"]


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(job):
    src, obj, verbose = job
    cmd = [NVCC, "-c", src, "-o", obj] + FLAGS + (["-Xptxas", "-v"] if verbose else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), (r.stdout + r.stderr)[-6000:]))
    return r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr_time = _deps_mtime()
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s).replace(".cu", ".o"))
        objs.append(o)
        if force or verbose or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            jobs.append((s, o, verbose))
    logs = []
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(_compile, jobs))
    stale = [o for o in glob.glob(os.path.join(OBJ, "*.o")) if o not in objs]
    for o in stale:
        os.remove(o)
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), (r.stdout + r.stderr)[-6000:]))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
