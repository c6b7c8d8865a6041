import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The reference ThreadPool is a process-wide singleton sized on first use
# (threading/threadpool.cpp:30-34); pin it so oracle results do not depend on the host. ONE worker
# thread: with more, the reference's threaded stages race on std::vector<bool> validity flags
# (flipsolver2d.cpp:1159,1373-1375: bits of one 64-bit word written from several ranges), which makes
# its trajectories differ from run to run at the 1e-4 level. The thread-count dependent convergence
# test (vmath.cpp:100-136) is exercised for T = 8 in a subprocess by test_pcg_gpu.py.
ORACLE_THREADS = 1
os.environ.setdefault("FS2D_ORACLE_THREADS", str(ORACLE_THREADS))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from flipsolver2d_b200 import capi
        return capi.lib().fs2d_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("scenes")


@pytest.fixture(scope="session")
def ref_mod():
    from oracle import ref
    if not ref.available(strict=True):
        pytest.skip("oracle/_ref not built")
    ref.load(strict=True, threads=ORACLE_THREADS)
    return ref
