"""Implicit viscosity against golden vectors (tests/golden/viscosity_systems.npz, made by tools/make_golden_viscosity.py
from the reference's own assembly): the numpy / scipy restatements of both systems -- the ones the GPU tests use to
predict iteration counts -- reproduce the stored outputs on CPU. Light model: the Upper view of LightViscosityModel's
matrix; heavy model: HeavyViscosityModel's coupled U+V system (viscositymodel.cpp:202-398)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("stages_gpu_helpers", os.path.join(HERE, "test_stages_gpu.py"))
stages = importlib.util.module_from_spec(spec)
spec.loader.exec_module(stages)
G = np.load(os.path.join(HERE, "golden", "viscosity_systems.npz"))


def test_light_model_restatement_matches_golden():
    I, J = int(G["light_I"]), int(G["light_J"])
    mat, mu = G["light_material"].reshape(I, J), G["light_viscosity"].reshape(I, J)
    dt, rho = float(G["light_dt"]), float(G["light_density"])
    u0, v0, u1, v1 = G["light_u0"], G["light_v0"], G["light_u1"], G["light_v1"]
    xu, it_u = stages._viscosity_cg_numpy(mat, mu, u0[:I * J].reshape(I, J), dt, rho)
    xv, it_v = stages._viscosity_cg_numpy(mat, mu, v0.reshape(I, J + 1)[:, :J], dt, rho)
    assert it_u > 0 and it_v > 0
    assert np.linalg.norm(xu.ravel() - u1[:I * J]) / np.linalg.norm(u1[:I * J]) < 1e-6
    assert np.linalg.norm(xv - v1.reshape(I, J + 1)[:, :J]) / np.linalg.norm(v1) < 1e-6
    assert np.array_equal(u1[I * J:], u0[I * J:])  # U's last row is outside the system
    assert np.linalg.norm(u1 - u0) / np.linalg.norm(u0) > 1e-3


def test_heavy_model_restatement_matches_golden_bit_for_bit():
    I, J = int(G["heavy_I"]), int(G["heavy_J"])
    un, vn, it = stages._heavy_viscosity_numpy(G["heavy_viscosity"].reshape(I, J), G["heavy_u0"], G["heavy_v0"], float(G["heavy_dt"]),
                                               float(G["heavy_dx"]), float(G["heavy_density"]))
    assert it > 0
    assert np.array_equal(un, G["heavy_u1"]) and np.array_equal(vn, G["heavy_v1"])
