"""The C-ABI libraries load and export every symbol their headers declare; without a GPU the product
path refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from flipsolver2d_b200 import capi, host_api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fs2d_exports_every_declared_symbol():
    L = capi.lib()
    names = capi.header_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(names) == set(L._fs2d_signatures), set(names) ^ set(L._fs2d_signatures)


def test_fs2d_host_exports_every_declared_symbol():
    L = host_api.lib()
    missing = [n for n in host_api.header_symbols() if not hasattr(L, n)]
    assert not missing, missing


def test_headers_have_no_torch_types():
    for h in ("fs2d.h", "fs2d_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments may mention torch, signatures may not
        assert not re.search(r"torch|at::|Tensor", code)


def test_params_struct_layout_matches_header():
    # 16 int32 + 3 double + 14 float, no padding surprises between the ctypes mirror and the C struct
    assert C.sizeof(capi.Params) == 16 * 4 + 3 * 8 + 14 * 4
    assert C.sizeof(capi.Source) == 8 * 4


def test_product_path_fails_loudly_without_gpu(tmp_path):
    if capi.lib().fs2d_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.Fs2dError):
        capi.Device(16, 16)
    s = host_api.Solver(scenes.write_scene(scenes.dam_break(32, "flip"), str(tmp_path / "s.json")))
    s.prepare_host()          # host-only work is fine
    with pytest.raises(capi.Fs2dError):
        s.step_frame()        # stepping needs the device


def test_product_package_never_imports_the_oracle():
    """Nothing under flipsolver2d_b200/ imports, links or opens anything under oracle/."""
    pkg = os.path.join(ROOT, "flipsolver2d_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|oracle/|oracle\\.|libfs2d_ref|ref_api\.h|ref_harness")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), os.path.join(dirpath, f)
