"""Shared helpers for the parity tests: build a reference solver (oracle/_ref) for a scene,
mirror its complete state into an fs2d device handle, canonical particle ordering."""
import numpy as np

from flipsolver2d_b200 import capi, scenes

SIM_OF = {"flip": capi.SIM_LIQUID, "fluid": capi.SIM_LIQUID, "smoke": capi.SIM_SMOKE, "fire": capi.SIM_FIRE,
          "nbflip": capi.SIM_NBFLIP}
HANDLING_OF = {"particle": capi.PARAMS_PARTICLE, "hybrid": capi.PARAMS_HYBRID, "grid": capi.PARAMS_GRID}

STATE_GRIDS = ["U", "V", "U_VALID", "V_VALID", "SAVED_U", "SAVED_V", "MATERIAL", "FLUID_SDF", "SOLID_SDF", "VISCOSITY",
               "DENSITY", "COUNTS", "EMITTER_ID", "SOLID_ID", "DIVERGENCE_CONTROL", "TEST", "KNOWN_CENTERED"]
SMOKE_GRIDS = ["TEMPERATURE", "CONCENTRATION"]


def scene_tables(scene):
    """Obstacle friction list and source table in the order JsonSceneReader adds them
    (Utils/jsonscenereader.cpp:187-287)."""
    friction, sources = [], []
    for o in scene["solver"]["objects"]:
        if not o.get("enabled", True):
            continue
        if o["type"] == "solid":
            friction.append(float(int(o.get("friction", 0))))  # tryGetValue(..., "friction", 0) reads an int (jsonscenereader.cpp:228)
        elif o["type"] == "source":
            vel = o.get("velocity", [0.0, 0.0])
            sources.append(dict(viscosity=float(o.get("viscosity", 0.0)), temperature=float(o.get("temperature", 273.0)),
                                concentration=float(o.get("concentration", 1.0)), fuel=float(o.get("fuel", 1.0)),
                                divergence=float(o.get("divergence", 0.0)), velocity_x=float(vel[0]),
                                velocity_y=float(vel[1]), transfer_velocity=int(bool(o.get("transferVelocity", False)))))
    return friction, sources


def property_layout(sim):
    """Column order the reference creates (flipsolver2d.cpp:72,1613-1616; flipsmokesolver.cpp:570-574;
    flipfiresolver.cpp:180-185): testValue first, then the solver's own columns."""
    if sim in ("flip", "fluid", "nbflip"):
        return dict(num_properties=2, test_property=0, viscosity_property=1, temperature_property=-1,
                    concentration_property=-1, fuel_property=-1)
    if sim == "smoke":
        return dict(num_properties=3, test_property=0, viscosity_property=-1, concentration_property=1,
                    temperature_property=2, fuel_property=-1)
    return dict(num_properties=4, test_property=0, viscosity_property=-1, concentration_property=1,
                temperature_property=2, fuel_property=3)


def make_ref(ref_mod, scene, path, frames=0, strict=True):
    scenes.write_scene(scene, str(path))
    s = ref_mod.RefSolver(str(path), strict=strict)
    for _ in range(frames):
        s.step_frame()
    return s


def make_device(s, scene, conv_threads=0, **over):
    p = s.params()
    st = scene["settings"]
    sim = st["simType"]
    kw = dict(dx=p["dx"], fluid_density=p["fluidDensity"], pcg_iter_limit=int(p["pcgIterLimit"]),
              particles_per_cell=int(p["ppc"]), sim_type=SIM_OF[sim],
              parameter_handling=HANDLING_OF[st.get("parameterHandlingMethod", "particle")],
              viscosity_enabled=int(p["viscosityEnabled"]), convergence_threads=conv_threads,
              project_tolerance=p["projectTolerance"], gravity_x=p["gx"], gravity_y=p["gy"], pic_ratio=p["picRatio"],
              particle_scale=p["particleScale"],
              ambient_temperature=float(st.get("ambientTemperature", 273.0)),
              temperature_decay=float(st.get("temperatureDecayRate", 0.0)),
              concentration_decay=float(st.get("concentrationDecayRate", 0.0)),
              buoyancy_factor=float(st.get("buoyancyFactor", 1.0)), soot_factor=float(st.get("sootFactor", 1.0)))
    kw.update(property_layout(sim))
    kw.update(over)
    d = capi.Device(s.I, s.J, **kw)
    friction, sources = scene_tables(scene)
    d.set_obstacles(friction)
    d.set_sources(sources)
    return d


def sync_state(s, d, sim="flip"):
    """Copy every grid and all particles of the reference solver into the device handle."""
    for g in STATE_GRIDS:
        d.upload(g, s.grid(g))
    if sim in ("smoke", "fire"):
        for g in SMOKE_GRIDS:
            d.upload(g, s.grid(g))
    if sim == "fire":
        d.upload("FUEL", s.grid("FUEL"))
    pos, vel, props, bins = s.particles()
    d.upload_particles(pos, vel, props)
    if len(pos):
        d.set_storage_bins(bins)  # the reference may hold particles in a bin that is not their position's
    d.set_step_dt(s.params()["stepDt"])


def canonical(pos, vel=None, props=None, J=None):
    """Order particles by (cell, x, y) -- the device's order -- so two sets can be compared."""
    i = np.floor(pos[:, 0]).astype(np.int64)
    j = np.floor(pos[:, 1]).astype(np.int64)
    order = np.lexsort((pos[:, 1], pos[:, 0], i * J + j))
    out = [pos[order]]
    if vel is not None:
        out.append(vel[order])
    if props is not None:
        out.append(props[:, order])
    return out if len(out) > 1 else out[0]


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (n if n > 0 else 1.0))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if len(a) else 0.0
