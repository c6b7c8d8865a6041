#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- builds the reference oracle into oracle/_ref/.

Compiles the UNMODIFIED ArtNlk/FlipSolver2d sources from where they lie under
/root/reference (FlipSolver2dLib/*.cpp minus the *_sse42.cpp files that are only built
with FLUID_ACCEL=sse, FlipSolver2dLib/CMakeLists.txt:60-90; threading/threadpool.cpp;
Utils/jsonscenereader.cpp) together with oracle/ref_harness.cpp, against
oracle/shim (Eigen stand-in) and the nlohmann/json 3.11.3 header that ships in the
image (cudnn_frontend/thirdparty). No reference source is copied; only objects and the
two shared libraries are written, all under oracle/_ref/ (git-ignored, gpurun-shipped).

Two variants:
  libfs2d_ref.so         the reference's Release flags (CMakeLists.txt:27,44-48,59-62,80:
                         -O3 -mavx2 -ffast-math -fopenmp -DFLUID_AVX2 -DNUMPY_LOGGING) --
                         the timing baseline. -march=native is replaced by
                         -march=x86-64-v3 because the library is built in a CPU-only
                         container and executed on a different host (the B200 box).
  libfs2d_ref_strict.so  FAST_MATH=OFF plus -ffp-contract=off: IEEE-exact float
                         arithmetic, used for parity so that 1-ulp decisions
                         (sdf < 0, weight thresholds) are reproducible.

The reference's own build system (cmake + FetchContent downloads) is not run.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

REF = os.environ.get("FS2D_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
JSON_INC_CANDIDATES = [
    os.path.join(sys.prefix, "lib/python3.12/site-packages/include/cudnn_frontend/thirdparty"),
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/include/cudnn_frontend/thirdparty",
]

COMMON = ["-std=c++20", "-fPIC", "-fopenmp", "-DFLUID_AVX2", "-DNUMPY_LOGGING", "-w"]
VARIANTS = {
    "libfs2d_ref.so": ["-O3", "-march=x86-64-v3", "-mavx2", "-ffast-math"],
    "libfs2d_ref_strict.so": ["-O2", "-march=x86-64-v3", "-mavx2", "-ffp-contract=off"],
}


def reference_sources():
    lib = sorted(glob.glob(os.path.join(REF, "FlipSolver2dLib", "*.cpp")))
    lib = [s for s in lib if not s.endswith("_sse42.cpp")]
    lib.append(os.path.join(REF, "FlipSolver2dLib", "threading", "threadpool.cpp"))
    lib.append(os.path.join(REF, "Utils", "jsonscenereader.cpp"))
    return lib


def json_include():
    for c in JSON_INC_CANDIDATES:
        if os.path.exists(os.path.join(c, "nlohmann", "json.hpp")):
            return c
    raise RuntimeError("nlohmann/json.hpp not found in the image")


def compile_one(args):
    src, obj, flags = args
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj
    cmd = ["g++", "-c", src, "-o", obj] + flags
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return obj


def build(force=False):
    """Build both variants. Returns the list of library paths."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present (prebuilt oracle/_ref is used on the GPU box)" % REF)
    os.makedirs(OUT, exist_ok=True)
    inc = [
        "-I" + os.path.join(HERE, "shim"),
        "-I" + json_include(),
        "-I" + os.path.join(REF, "FlipSolver2dLib"),
        "-I" + os.path.join(REF, "FlipSolver2dLib", "threading"),
        "-I" + os.path.join(REF, "Utils"),
        "-I" + HERE,
    ]
    srcs = reference_sources() + [os.path.join(HERE, "ref_harness.cpp")]
    libs = []
    for libname, vflags in VARIANTS.items():
        tag = libname.replace("libfs2d_", "").replace(".so", "")
        objdir = os.path.join(OUT, "obj_" + tag)
        os.makedirs(objdir, exist_ok=True)
        jobs = []
        for s in srcs:
            obj = os.path.join(objdir, os.path.basename(s).replace(".cpp", ".o"))
            if force and os.path.exists(obj):
                os.remove(obj)
            jobs.append((s, obj, COMMON + vflags + inc))
        # the harness depends on its own header too
        hobj = jobs[-1][1]
        if os.path.exists(hobj) and os.path.getmtime(hobj) < os.path.getmtime(os.path.join(HERE, "ref_api.h")):
            os.remove(hobj)
        with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(compile_one, jobs))
        lib = os.path.join(OUT, libname)
        newest = max(os.path.getmtime(o) for o in objs)
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest:
            cmd = ["g++", "-shared", "-o", lib] + objs + ["-fopenmp", "-lpthread", "-Wl,-Bsymbolic-functions"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        libs.append(lib)
    return libs


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print("built", p)
