import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The reference ThreadPool is a process-wide singleton sized on first use
# (threading/threadpool.cpp:30-34); pin it so oracle results do not depend on the host.
ORACLE_THREADS = 8
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
