#include "jsonscenereader.h"

#include <fstream>
#include <iostream>

std::shared_ptr<FlipSolver> JsonSceneReader::loadJson(std::string fileName)
{
    try
    {
        std::ifstream file(fileName);
        if (!file.is_open())
        {
            std::cout << "errorOpening scene file " << fileName << "\n";
            return std::shared_ptr<FlipSolver>();
        }
        json scene;
        file >> scene;
        json settings = scene["settings"];
        std::shared_ptr<FlipSolver> solver;
        switch (simMethodFromName(settings["simType"].get<std::string>()))
        {
        case SIMULATION_LIQUID:
        {
            FlipSolverParameters p;
            populateFlipSolverParamsFromJson(&p, settings);
            solver = std::make_shared<FlipSolver>(&p);
            break;
        }
        case SIMULATION_SMOKE:
        {
            SmokeSolverParameters p;
            populateFlipSolverParamsFromJson(&p, settings);
            populateSmokeSolverParamsFromJson(&p, settings);
            solver = std::make_shared<FlipSmokeSolver>(&p);
            break;
        }
        case SIMULATION_FIRE:
        {
            FireSolverParameters p;
            populateFlipSolverParamsFromJson(&p, settings);
            populateFireSolverParamsFromJson(&p, settings);
            solver = std::make_shared<FlipFireSolver>(&p);
            break;
        }
        case SIMULATION_NBFLIP:
        {
            NBFlipParameters p;
            populateFlipSolverParamsFromJson(&p, settings);
            populateNBFlipSolverParamsFromJson(&p, settings);
            solver = std::make_shared<NBFlipSolver>(&p);
            break;
        }
        }
        solver->initAdditionalParameters();
        objectsFromJson(scene["solver"], solver);
        return solver;
    }
    catch (std::exception &e)
    {
        std::cout << e.what();
        return std::shared_ptr<FlipSolver>();
    }
}

// Keys and defaults of jsonscenereader.cpp:79-136. Value types matter: tryGetValue<T> parses the JSON
// number as the type of the default (e.g. "picRatio" as double, "friction" as int).
void JsonSceneReader::populateFlipSolverParamsFromJson(FlipSolverParameters *p, json s)
{
    const float scale = tryGetValue(s, "scale", 1.f);
    p->fluidDensity = tryGetValue(s, "density", 1.f);
    p->seed = tryGetValue(s, "seed", 0);
    p->particlesPerCell = s["particlesPerCell"].get<int>();
    std::pair<float, float> g = tryGetValue(s, "globalAcceleration", std::pair(9.8, 0.f));
    p->globalAcceleration = g;
    p->resolution = s["resolution"].get<int>();
    p->fps = s["fps"].get<int>();
    p->maxSubsteps = tryGetValue(s, "maxSubsteps", 30);
    p->picRatio = tryGetValue(s, "picRatio", 0.03);
    p->cflNumber = tryGetValue(s, "cflNumber", 10.f);
    p->particleScale = tryGetValue(s, "particleScale", 0.8);
    p->pcgIterLimit = tryGetValue(s, "pcgIterLimit", 200);
    p->viscosityEnabled = tryGetValue(s, "viscosityEnabled", false);
    p->domainSizeI = s["domainSizeI"].get<float>() * scale;
    p->domainSizeJ = s["domainSizeJ"].get<float>() * scale;
    p->useHeavyViscosity = tryGetValue(s, "heavyViscosity", false);
    p->sceneScale = scale;
    // grid from resolution and domain aspect (jsonscenereader.cpp:103-118)
    if (p->domainSizeI > p->domainSizeJ)
    {
        p->dx = static_cast<float>(p->domainSizeI) / p->resolution;
        p->gridSizeI = p->resolution;
        p->gridSizeJ = (static_cast<float>(p->domainSizeJ) / static_cast<float>(p->domainSizeI)) * p->resolution;
    }
    else
    {
        p->dx = static_cast<float>(p->domainSizeJ) / p->resolution;
        p->gridSizeJ = p->resolution;
        p->gridSizeI = (static_cast<float>(p->domainSizeI) / static_cast<float>(p->domainSizeJ)) * p->resolution;
    }
    const std::string handling = tryGetValue(s, "parameterHandlingMethod", std::string("particle"));
    p->parameterHandlingMethod = PARTICLE;
    if (handling == "hybrid") p->parameterHandlingMethod = HYBRID;
    if (handling == "grid") p->parameterHandlingMethod = GRID;
    p->simulationMethod = simMethodFromName(s["simType"].get<std::string>());
}

void JsonSceneReader::populateNBFlipSolverParamsFromJson(NBFlipParameters *, json) {}

void JsonSceneReader::populateSmokeSolverParamsFromJson(SmokeSolverParameters *p, json s)
{
    p->ambientTemperature = tryGetValue(s, "ambientTemperature", 273.0f);
    p->temperatureDecayRate = tryGetValue(s, "temperatureDecayRate", 0.0);
    p->concentrationDecayRate = tryGetValue(s, "concentrationDecayRate", 0.0);
    p->buoyancyFactor = tryGetValue(s, "buoyancyFactor", 1.0);
    p->sootFactor = tryGetValue(s, "sootFactor", 1.0);
}

void JsonSceneReader::populateFireSolverParamsFromJson(FireSolverParameters *p, json s)
{
    populateSmokeSolverParamsFromJson(p, s);
    p->ignitionTemperature = tryGetValue(s, "ignitionTemp", 250.f);
    p->burnRate = tryGetValue(s, "burnRate", 0.05f);
    p->smokeProportion = tryGetValue(s, "smokeEmission", 1.f);
    p->heatProportion = tryGetValue(s, "heatEmission", 1.f);
    p->divergenceProportion = tryGetValue(s, "billowing", 0.1f);
}

SimulationMethod JsonSceneReader::simMethodFromName(const std::string &name)
{
    if (name == "smoke") return SIMULATION_SMOKE;
    if (name == "fire") return SIMULATION_FIRE;
    if (name == "nbflip") return SIMULATION_NBFLIP;
    return SIMULATION_LIQUID;  // "fluid", "flip" and anything unknown
}

void JsonSceneReader::objectsFromJson(json solverJson, std::shared_ptr<FlipSolver> solver)
{
    std::vector<json> objects = solverJson["objects"].get<std::vector<json>>();
    for (json &o : objects) addObjectFromJson(o, solver);
}

static Geometry2d polygonFromJson(nlohmann::json &j, float sceneScale)
{
    std::vector<std::pair<float, float>> verts = j["verts"].get<std::vector<std::pair<float, float>>>();
    Geometry2d geo;
    for (auto &v : verts) geo.addVertex(sceneScale * Vec3(v.first, v.second));
    return geo;
}

Emitter JsonSceneReader::emitterFromJson(json j, float sceneScale)
{
    Geometry2d geo = polygonFromJson(j, sceneScale);
    Emitter e(geo);
    e.setTemperature(tryGetValue(j, "temperature", 273.f));
    e.setConcentrartion(tryGetValue(j, "concentration", 1.f));
    e.setViscosity(tryGetValue(j, "viscosity", 0.f));
    e.setFuel(tryGetValue(j, "fuel", 1.f));
    e.setDivergence(tryGetValue(j, "divergence", 0.f));
    std::pair<float, float> vel = tryGetValue(j, "velocity", std::pair<float, float>(0.f, 0.f));
    e.setVelocity(Vec3(vel));
    e.setVelocityTransfer(tryGetValue(j, "transferVelocity", false));
    return e;
}

Obstacle JsonSceneReader::obstacleFromJson(json j, float sceneScale)
{
    const float friction = tryGetValue(j, "friction", 0);  // parsed as int, like the reference (:228)
    Geometry2d geo = polygonFromJson(j, sceneScale);
    return Obstacle(friction, geo);
}

Sink JsonSceneReader::sinkFromJson(json j, float sceneScale)
{
    Geometry2d geo = polygonFromJson(j, sceneScale);
    return Sink(tryGetValue(j, "divergence", 0.f), geo);
}

void JsonSceneReader::addObjectFromJson(json j, std::shared_ptr<FlipSolver> solver)
{
    const std::string type = j["type"].get<std::string>();
    const float scale = solver->sceneScale();
    if (!tryGetValue(j, "enabled", true)) return;
    if (type == "solid")
    {
        Obstacle o = obstacleFromJson(j, scale);
        solver->addGeometry(o);
    }
    else if (type == "source")
    {
        Emitter e = emitterFromJson(j, scale);
        solver->addSource(e);
    }
    else if (type == "sink")
    {
        Sink s = sinkFromJson(j, scale);
        solver->addSink(s);
    }
    else if (type == "fluid")
    {
        Emitter e = emitterFromJson(j, scale);
        solver->addInitialFluid(e);
    }
}
