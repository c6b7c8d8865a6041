// Autobench: headless scene benchmark with the reference's command line
// (AutoBench/benchmarkrunnerapplication.cpp:14-48): -i <scene.json | directory of *.json>
// -o <output directory> -s <frames per scene, default 60>. For every scene it loads the JSON through
// JsonSceneReader, calls stepFrame() and records SolverStats per frame, then writes <output>/Stats.xlsx
// -- one worksheet per scene, the 18 columns of AutoBench/benchruntable.h:28-49 (benchruntable.h here:
// a dependency-free xlsx writer, OpenXLSX is not available) -- and the same table as Stats.csv.
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../benchruntable.h"
#include "../jsonscenereader.h"

namespace fs = std::filesystem;

static std::string option(int argc, char **argv, const std::string &name)
{
    for (int k = 1; k + 1 < argc; k++)
        if (name == argv[k]) return argv[k + 1];
    return "";
}

struct SceneRun
{
    std::string name;
    std::vector<SolverStats> frames;
    size_t particles = 0;
    size_t cells = 0;
};

static bool runScene(const fs::path &scene, int frames, std::vector<SceneRun> &out)
{
    std::shared_ptr<FlipSolver> solver = JsonSceneReader::loadJson(scene.string());
    if (!solver)
    {
        std::cout << "Failed to load scene: " << scene.string() << '\n';
        return false;
    }
    std::cout << "Starting scene: " << scene.string() << '\n';
    SceneRun run;
    run.name = scene.stem().string();
    for (int f = 0; f < frames; f++)
    {
        solver->stepFrame();
        run.frames.push_back(solver->timeStats());
    }
    run.particles = solver->particleCount();
    run.cells = solver->cellCount();
    std::cout << "Finished scene: " << scene.string() << '\n';
    out.push_back(run);
    return true;
}

int main(int argc, char **argv)
{
    fs::path input = option(argc, argv, "-i");
    fs::path output = option(argc, argv, "-o");
    const std::string steps = option(argc, argv, "-s");
    if (input.empty()) input = fs::current_path();
    if (output.empty()) output = fs::current_path();
    const int frames = steps.empty() ? 60 : std::stoi(steps);

    std::vector<SceneRun> runs;
    try
    {
        if (!fs::is_directory(input) || input.extension() == ".json")
        {
            runScene(input, frames, runs);
        }
        else
        {
            for (const fs::directory_entry &e : fs::directory_iterator(input))
                if (!fs::is_directory(e) && e.path().extension() == ".json") runScene(e.path(), frames, runs);
        }
    }
    catch (std::exception &e)
    {
        std::cerr << "Autobench: " << e.what() << std::endl;
        return 1;
    }

    BenchRunTable table;
    table.setOutputFile(output / "Stats.xlsx");
    for (const SceneRun &run : runs)
    {
        for (const SolverStats &st : run.frames) table.addStepTiming(st);
        table.finishScene(run.name);
    }
    if (!table.save()) std::cerr << "Autobench: cannot write " << (output / "Stats.xlsx").string() << std::endl;

    std::ofstream csv(output / "Stats.csv");
    for (const SceneRun &run : runs)
    {
        double totalMs = 0.0;
        long substeps = 0;
        csv << "# scene," << run.name << ",cells," << run.cells << ",particles," << run.particles << "\n";
        for (int c = 0; c < 18; c++) csv << BenchRunTable::getColumnHeader(c) << (c + 1 < 18 ? "," : "\n");
        for (size_t f = 0; f < run.frames.size(); f++)
        {
            const SolverStats &s = run.frames[f];
            const SolverStats::StageTimings t = s.timings();
            csv << f + 1 << "," << s.substepCount() << "," << s.frameTime();
            for (float v : t) csv << "," << v;
            csv << "," << s.pressureIterations() << "," << s.densityIterations() << "," << s.viscosityIterations() << "\n";
            totalMs += s.frameTime();
            substeps += s.substepCount();
        }
        std::cout << run.name << ": " << run.frames.size() << " frames, " << substeps << " substeps, " << totalMs << " ms, "
                  << (totalMs > 0 ? substeps / (totalMs * 1e-3) : 0.0) << " substeps/s\n";
    }
    return 0;
}
