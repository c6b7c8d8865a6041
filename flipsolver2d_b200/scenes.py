"""Scene JSON builders for the benchmark configurations (BASELINE.json `configs`).

Scenes use the reference's scene format (README.md:47-135, parsed by
Utils/jsonscenereader.cpp:79-287): a `settings` object and `solver.objects`, each object a
polygon in domain units. The geometry below restates the two scenes the benchmark
configs name -- Liquid2dRender/scenes/dam_break.json (50x50 tank, four 3-unit walls, a
22x10 fluid block) and smoke_test.json (tank + wedge obstacle + sink + hot source) --
with the resolution / sim type / solver switches as parameters.
"""
import json
import os


def _rect(i0, j0, i1, j1):
    return [[i0, j0], [i0, j1], [i1, j1], [i1, j0]]


def tank_walls(size_i=50, size_j=50, t=3):
    return [
        {"type": "solid", "enabled": True, "verts": [[size_i - t, 0], [size_i - t, size_j], [size_i, size_j], [size_i, 0]]},
        {"type": "solid", "enabled": True, "verts": [[0, 0], [0, t], [size_i, t], [size_i, 0]]},
        {"type": "solid", "enabled": True, "verts": [[0, 0], [0, size_j], [t, size_j], [t, 0]]},
        {"type": "solid", "enabled": True, "verts": [[0, size_j - t], [0, size_j], [size_i, size_j], [size_i, size_j - t]]},
    ]


def dam_break(resolution=128, sim_type="flip", ppc=8, pcg_iter_limit=None, viscosity_enabled=False,
              pic_ratio=0.03, seed=0, dense_fill=False, max_substeps=10, fluid_viscosity=10):
    """SURVEY.md 8(d) configs 1, 2, 4, 5. `dense_fill` is the near-full tank variant."""
    settings = {
        "simType": sim_type,
        "domainSizeI": 50,
        "domainSizeJ": 50,
        "resolution": int(resolution),
        "fps": 30,
        "density": 0.5,
        "maxSubsteps": int(max_substeps),
        "seed": int(seed),
        "particlesPerCell": int(ppc),
        "particleScale": 0.8,
        "picRatio": pic_ratio,
        "cflNumber": 5,
        "viscosityEnabled": bool(viscosity_enabled),
        "globalAcceleration": [9.8, 0],
    }
    if pcg_iter_limit is not None:
        settings["pcgIterLimit"] = int(pcg_iter_limit)
    fluid = _rect(5, 3, 47, 47) if dense_fill else [[25, 3], [25, 13], [47, 13], [47, 3]]
    objects = tank_walls() + [{"type": "fluid", "viscosity": fluid_viscosity, "enabled": True, "verts": fluid}]
    return {"settings": settings, "solver": {"objects": objects}}


def smoke_test(resolution=256, ppc=4, parameter_handling="particle", pcg_iter_limit=None, sim_type="smoke"):
    """SURVEY.md 8(d) config 3 (grid-advected temperature/soot when parameter_handling="grid")."""
    settings = {
        "simType": sim_type,
        "domainSizeI": 50,
        "domainSizeJ": 50,
        "resolution": int(resolution),
        "fps": 30,
        "maxSubsteps": 10,
        "seed": 0,
        "particlesPerCell": int(ppc),
        "picRatio": 0.03,
        "cflNumber": 5,
        "temperatureDecayRate": 0.3,
        "concentrationDecayRate": 0.05,
        "buoyancyFactor": 1,
        "sootFactor": 0.5,
        "globalAcceleration": [9.8, 0],
        "parameterHandlingMethod": parameter_handling,
    }
    if pcg_iter_limit is not None:
        settings["pcgIterLimit"] = int(pcg_iter_limit)
    if sim_type == "fire":
        settings.update({"burnRate": 0.8, "ignitionTemp": 250, "smokeEmission": 1, "heatEmission": 500,
                         "billowing": 10, "density": 0.01})
    objects = [
        {"type": "solid", "verts": [[47, 0], [47, 50], [50, 50], [50, 0]]},
        {"type": "solid", "verts": [[3, 0], [3, 3], [47, 3], [47, 0]]},
        {"type": "solid", "verts": [[0, 0], [0, 50], [3, 50], [3, 0]]},
        {"type": "solid", "verts": [[3, 47], [3, 50], [47, 50], [47, 47]]},
        {"type": "solid", "verts": [[30, 15], [30, 35], [35, 25]]},
        {"type": "sink", "divergence": -10, "verts": [[3, 40], [3, 47], [8, 47], [8, 40]]},
        {"type": "source", "viscosity": 0, "divergence": 0, "temperature": 1273, "concentration": 1,
         "fuel": 1, "velocity": [-3, 0], "verts": [[40, 20], [40, 30], [46, 30], [46, 20]]},
    ]
    return {"settings": settings, "solver": {"objects": objects}}


def source_sink(resolution=64, sim_type="flip", ppc=8, transfer_velocity=True):
    """Small scene with an emitter and a sink so reseeding / death paths are exercised."""
    scene = dam_break(resolution, sim_type, ppc)
    scene["solver"]["objects"] += [
        {"type": "source", "viscosity": 0.2, "velocity": [4, 1], "transferVelocity": bool(transfer_velocity),
         "verts": _rect(8, 30, 12, 36)},
        {"type": "sink", "divergence": 0, "verts": _rect(42, 40, 47, 47)},
        {"type": "solid", "friction": 0.3, "verts": [[30, 20], [36, 30], [40, 20]]},
    ]
    return scene


def write_scene(scene, path):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(scene, f, indent=1)
    return path
