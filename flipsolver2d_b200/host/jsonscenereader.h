// Host mirror of Utils/jsonscenereader.h: scene JSON -> parameters -> solver object -> scene objects.
// Same entry point and the same JSON keys / defaults as the reference reader
// (Utils/jsonscenereader.cpp:8-293), so the reference's scene files load unchanged.
#ifndef FS2D_HOST_JSONSCENEREADER_H
#define FS2D_HOST_JSONSCENEREADER_H

#include <memory>
#include <string>

#include <nlohmann/json.hpp>

#include "solvers.h"

class JsonSceneReader
{
public:
    JsonSceneReader() = default;
    // Returns an empty pointer on any error after printing the message (jsonscenereader.cpp:68-74).
    static std::shared_ptr<FlipSolver> loadJson(std::string fileName);

protected:
    using json = nlohmann::json;
    static void populateFlipSolverParamsFromJson(FlipSolverParameters *p, json settingsJson);
    static void populateNBFlipSolverParamsFromJson(NBFlipParameters *p, json settingsJson);
    static void populateSmokeSolverParamsFromJson(SmokeSolverParameters *p, json settingsJson);
    static void populateFireSolverParamsFromJson(FireSolverParameters *p, json settingsJson);
    static SimulationMethod simMethodFromName(const std::string &name);
    static void objectsFromJson(json solverJson, std::shared_ptr<FlipSolver> solver);
    static Emitter emitterFromJson(json emitterJson, float sceneScale);
    static Obstacle obstacleFromJson(json obstacleJson, float sceneScale);
    static Sink sinkFromJson(json sinkJson, float sceneScale);
    static void addObjectFromJson(json objectJson, std::shared_ptr<FlipSolver> solver);
    template <class T> static T tryGetValue(json input, std::string key, T defaultValue)
    {
        return input.contains(key) ? input[key].get<T>() : defaultValue;
    }
};

#endif
