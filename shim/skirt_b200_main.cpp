// skirt_b200 -- SKIRT 9 with the photon life cycle on a B200: the reference's own object model, .ski reader, setup code,
// probes and output writers (linked unmodified from the reference's object files) driven by GpuLifeCycle, which runs
// every emission segment through libskirt9_b200.so (include/sk_engine.h).
//
//   skirt_b200 [-t threads] [-b] [-i indir] [-o outdir] [-g device[,device...]] [--host-setup] [--cpu] file.ski
//
// The structure follows SKIRT/main/SkirtMain.cpp:15-31 and SkirtCommandLineHandler::doSimulation
// (SKIRT/main/SkirtCommandLineHandler.cpp:295-400).  There is no CPU fallback: a configuration outside the accelerated
// path ends with a fatal error that names the reason (the reference's error convention, SkirtCommandLineHandler.cpp:372-400).
// `--cpu` is an explicit baseline mode: it runs the reference's own CPU life cycle (the same binary then is the reference)
// and says so in the log ("CPU life cycle (reference)"); it is never chosen on its own.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <typeinfo>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <valarray>
#include <vector>

#define private public
#define protected public
#include "MonteCarloSimulation.hpp"
#undef private
#undef protected

#include <cuda_runtime_api.h>

#include "BuildInfo.hpp"
#include "Console.hpp"
#include "FatalError.hpp"
#include "FileLog.hpp"
#include "FilePaths.hpp"
#include "GpuLifeCycle.hpp"
#include "ParallelFactory.hpp"
#include "ProcessManager.hpp"
#include "SignalHandler.hpp"
#include "SimulationItemRegistry.hpp"
#include "StringUtils.hpp"
#include "System.hpp"
#include "TimeLogger.hpp"
#include "XmlHierarchyCreator.hpp"
#include "XmlHierarchyWriter.hpp"

int main(int argc, char** argv)
{
    ProcessManager pm(&argc, &argv);
    System system(argc, argv);
    SignalHandler::InstallSignalHandlers();
    string version = BuildInfo::projectVersion();
    SimulationItemRegistry registry(version, "9");
    Console console;

    // ---- command line
    int threads = 0;
    std::vector<int> devices;
    bool brief = false, cpu = false, hostSetup = false;
    string inpath, outpath, skipath;
    for (int i = 1; i < argc; ++i)
    {
        string a = argv[i];
        auto value = [&]() -> string {
            if (i + 1 >= argc) throw FATALERROR("Missing value for option " + a);
            return argv[++i];
        };
        try
        {
            if (a == "-t")
                threads = std::stoi(value());
            else if (a == "-g")
            {
                // one CUDA device ordinal or a comma-separated list: the histories of every segment are split over them
                devices.clear();
                string list = value();
                size_t pos = 0;
                while (pos <= list.size())
                {
                    size_t comma = list.find(',', pos);
                    if (comma == string::npos) comma = list.size();
                    int device = std::stoi(list.substr(pos, comma - pos));
                    if (device < 0) throw FATALERROR("Negative CUDA device ordinal in -g " + list);
                    if (std::find(devices.begin(), devices.end(), device) != devices.end())
                        throw FATALERROR("CUDA device " + std::to_string(device) + " is listed twice in -g " + list);
                    devices.push_back(device);
                    pos = comma + 1;
                }
            }
            else if (a == "-i")
                inpath = value();
            else if (a == "-o")
                outpath = value();
            else if (a == "-b")
                brief = true;
            else if (a == "--cpu")
                cpu = true;
            else if (a == "--host-setup")
                hostSetup = true;
            else if (!a.empty() && a[0] != '-')
                skipath = a;
            else
                throw FATALERROR("Unknown option " + a);
        }
        catch (FatalError& error)
        {
            for (string line : error.message()) console.error(line);
            return EXIT_FAILURE;
        }
    }
    if (skipath.empty())
    {
        console.error("usage: skirt_b200 [-t threads] [-b] [-i indir] [-o outdir] [-g device[,device...]] [--host-setup] [--cpu] file.ski");
        return EXIT_FAILURE;
    }
    if (!StringUtils::endsWith(skipath, ".ski")) skipath += ".ski";

    // The CUDA context of the first device takes a few hundred milliseconds to create: start that now, on a thread of its
    // own, so that it overlaps with reading the ski file and building the simulation hierarchy instead of delaying the first
    // engine call (the octree construction during set-up).
    std::thread warmup;
    if (!cpu)
        warmup = std::thread([device = devices.empty() ? 0 : devices[0]]() {
            if (cudaSetDevice(device) == cudaSuccess) cudaFree(nullptr);
        });
    struct Joiner {
        std::thread& t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{warmup};

    string producer = "SKIRT " + version + " + B200 life-cycle engine (ABI " + std::to_string(sk_abi_version()) + ")";
    console.info("Welcome to " + producer);
    console.info("Constructing a simulation from ski file '" + skipath + "'...");
    try
    {
        auto schema = SimulationItemRegistry::getSchemaDef();
        auto topitem = XmlHierarchyCreator::readFile(schema, skipath);
        auto simulation = dynamic_cast<MonteCarloSimulation*>(topitem.get());
        if (!simulation) throw FATALERROR("The ski file does not hold a MonteCarloSimulation");

        simulation->filePaths()->setOutputPrefix(StringUtils::filenameBase(skipath));
        simulation->filePaths()->setInputPath(inpath);
        simulation->filePaths()->setOutputPath(outpath);
        if (threads > 0) simulation->parallelFactory()->setMaxThreadCount(threads);

        FileLog* log = new FileLog();
        simulation->log()->setLinkedLog(log);
        if (brief) simulation->log()->setLowestLevel(Log::Level::Success);
        XmlHierarchyWriter::write(simulation, schema, simulation->filePaths()->output("parameters.xml"), producer);
        log->setup();
        log->info(producer);

        try
        {
            // Simulation::setupAndRun (Simulation.cpp:12-35) with the run phase handed to the GPU driver
            simulation->_factory->setup();
            simulation->_log->setup();
            TimeLogger logger(simulation->_log, "simulation " + simulation->_paths->outputPrefix());
            // the octree of a DensityTreePolicy is constructed on the first device (--host-setup keeps the reference's loop)
            if (!cpu && !hostSetup && GpuLifeCycle::installDeviceTreeConstruction(simulation, devices.empty() ? 0 : devices[0]))
                simulation->_log->info("The spatial tree will be constructed on the GPU");
            simulation->setupSimulation();

            if (cpu)
            {
                // explicit baseline mode, never a fall-back
                simulation->_log->warning("CPU life cycle (reference): --cpu was given, the GPU engine is not used");
                simulation->runSimulation();
            }
            else
            {
                GpuLifeCycle gpu(simulation, devices);
                string why = gpu.unsupportedReason();
                if (!why.empty())
                    throw FATALERROR("This configuration is outside the GPU life cycle (" + why
                                     + "); skirt_b200 has no CPU fall-back (use the reference, or --cpu for its baseline mode)");
                {
                    TimeLogger uplog(simulation->_log, "GPU engine configuration");
                    gpu.configure();
                }
                gpu.runSimulation();
                const sk_counters_t& c = gpu.counters();
                simulation->_log->info("GPU life cycle: " + std::to_string(c.packets) + " packets, "
                                       + std::to_string(c.forward_segments + c.peel_segments) + " path segments, "
                                       + std::to_string(c.scatterings) + " scatterings, " + std::to_string(c.detections)
                                       + " detections, " + std::to_string(c.kernel_launches) + " kernel launches");
            }
        }
        catch (FatalError& error)
        {
            for (string line : error.message()) log->error(line, false);
            throw error;
        }
    }
    catch (FatalError& error)
    {
        for (string line : error.message()) console.error(line);
        return EXIT_FAILURE;
    }
    catch (const std::exception& except)
    {
        console.error("Standard Library Exception: " + string(except.what()));
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
