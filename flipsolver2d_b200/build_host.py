#!/usr/bin/env python3
"""Build the C++ host mirror of the reference's solver API into libfs2d_host.so (+ the Autobench
driver), in-tree. Compiled with -ffp-contract=off so that frame-0 rasterisation / seeding reproduce the
reference built without -ffast-math. Needs libfs2d_cuda.so (flipsolver2d_b200/build.py) to link."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libfs2d_host.so")
AUTOBENCH = os.path.join(HERE, "Autobench")
JSON_INC_CANDIDATES = [
    os.path.join(sys.prefix, "lib/python3.12/site-packages/include/cudnn_frontend/thirdparty"),
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/include/cudnn_frontend/thirdparty",
]
FLAGS = ["-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wno-sign-compare", "-pthread"]


def json_include():
    for c in JSON_INC_CANDIDATES:
        if os.path.exists(os.path.join(c, "nlohmann", "json.hpp")):
            return c
    raise RuntimeError("nlohmann/json.hpp not found in the image")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), (r.stdout + r.stderr)[-6000:]))


def _stale(target, deps):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(d) for d in deps)


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(HOST, "*.cpp")))
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    cuda_lib = os.path.join(HERE, "libfs2d_cuda.so")
    if not os.path.exists(cuda_lib):
        raise RuntimeError("libfs2d_cuda.so missing: run flipsolver2d_b200/build.py first")
    inc = ["-I" + json_include()]
    link = ["-L" + HERE, "-lfs2d_cuda", "-Wl,-rpath,$ORIGIN"]
    if force or _stale(LIB, srcs + hdrs + [cuda_lib]):
        _run(["g++", "-shared", "-o", LIB] + FLAGS + inc + srcs + link)
    main = os.path.join(HOST, "autobench", "main.cpp")
    if force or _stale(AUTOBENCH, [main, LIB]):
        _run(["g++", "-o", AUTOBENCH] + FLAGS + inc + [main, "-L" + HERE, "-lfs2d_host", "-lfs2d_cuda", "-Wl,-rpath,$ORIGIN"])
    return LIB, AUTOBENCH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
