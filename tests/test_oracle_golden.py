"""Pins the oracle (the unmodified reference compiled into oracle/_ref) against the only golden vectors
the reference ships: devDocs/matviz/{data,vin,vout}.txt, committed as
tests/golden/matviz_pressure_system.npz by tools/make_golden.py -- one real 64x64 dam-break pressure
system (402 fluid rows), its rhs and the pressure the reference's solver produced. CPU only."""
import os

import numpy as np
import pytest

from flipsolver2d_b200 import capi, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matviz_pressure_system.npz")


def golden_material(g):
    """The material grid behind the dump: 64x64 tank with the dam-break scene's 3-unit walls (4 cells),
    the dumped rows FLUID, everything else EMPTY. Every dumped diagonal is consistent with it."""
    I, J = (int(v) for v in g["size"])
    mat = np.full((I, J), capi.EMPTY, np.int8)
    mat[:4] = capi.SOLID
    mat[I - 4:] = capi.SOLID
    mat[:, :4] = capi.SOLID
    mat[:, J - 4:] = capi.SOLID
    m = mat.ravel().copy()
    m[g["index"]] = capi.FLUID
    return m


def numpy_operator(g, x):
    """A*x straight from the dumped rows (identity on non-fluid cells, pressuredata.h:227-236)."""
    I, J = (int(v) for v in g["size"])
    y = x.copy()
    idx = g["index"]
    xp = np.concatenate([np.zeros(J), x, np.zeros(J)])
    c = idx + J
    y[idx] = g["diag"] * xp[c] + g["i_neg"] * xp[c - J] + g["i_pos"] * xp[c + J] + g["j_neg"] * xp[c - 1] + g["j_pos"] * xp[c + 1]
    return y


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def golden_ref(ref_mod, scene_dir, golden):
    path = scenes.write_scene(scenes.dam_break(64, "flip"), str(scene_dir / "golden64.json"))
    s = ref_mod.RefSolver(path, strict=True)
    s.stage("FIRST_FRAME_INIT")
    s.set_grid("MATERIAL", golden_material(golden))
    s.set_step_dt(1.0 / 30.0)  # scale = dt/(rho dx^2) = 0.109227 for rho 0.5, dx 50/64
    s.stage("BUILD_MATRIX")
    yield s
    s.close()


def test_fixture_is_self_consistent(golden):
    """vout solves the dumped system for vin (to the 6 printed digits)."""
    r = numpy_operator(golden, golden["vout"]) - golden["vin"]
    assert np.abs(r).max() < 1e-4


def test_oracle_matrix_matches_golden_rows(golden, golden_ref):
    m = golden_ref.matrix()
    idx = golden["index"]
    assert abs(m["scale"] - 0.109227) < 1e-6
    assert np.array_equal(np.nonzero(m["is_unit"])[0], idx)
    assert np.abs(m["scale"] * m["count"][idx] - golden["diag"]).max() < 1e-6
    for bit, name in ((1, "i_neg"), (2, "i_pos"), (4, "j_neg"), (8, "j_pos")):
        assert np.abs(-m["scale"] * ((m["mask"][idx] & bit) != 0) - golden[name]).max() < 1e-6, name


def test_oracle_operator_and_solve_match_golden(golden, golden_ref):
    r = golden_ref.spmv(golden["vout"]) - golden["vin"]
    assert np.abs(r).max() < 1e-4
    x, iters = golden_ref.pcg(golden["vin"], 2000, 1e-6)
    assert iters < 2000
    assert np.abs(x - golden["vout"]).max() < 1e-3 * np.abs(golden["vout"]).max()
