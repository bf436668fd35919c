// GpuLifeCycle -- the reference-side shim of INTEGRATION.md made real: a C++ class that sits next to an UNMODIFIED
// SKIRT 9 MonteCarloSimulation object (built from the .ski file by the reference's own XmlHierarchyCreator and set up by
// the reference's own setup() code), extracts the flat tables of include/sk_engine.h from it, runs the photon life cycle
// of every emission segment on the GPU through the C ABI, and hands the tallies back to the reference's FluxRecorder and
// MediumSystem objects so that the reference's own probes and FITS / text writers produce the output files.
//
// It replaces exactly MonteCarloSimulation::runSimulation's use of performLifeCycle
// (SKIRT/core/MonteCarloSimulation.cpp:58-173, 335-403, 538-613); nothing of the reference is copied or modified: this
// file is compiled against the reference headers where they lie and linked with the reference's object files.
#ifndef GPULIFECYCLE_HPP
#define GPULIFECYCLE_HPP

#include <string>
#include <vector>
#include "sk_engine.h"

class MonteCarloSimulation;

class GpuLifeCycle
{
public:
    // one engine per listed CUDA device; with several devices every emission segment is split into interleaved blocks of
    // 16384 histories (device i runs every n-th block; one host thread per device) and the tallies are all-reduced with NCCL
    // exactly where the reference calls ProcessManager::sumToAll / sumToRoot (MediumSystem.cpp:1304-1313,
    // FluxRecorder.cpp:487-493)
    GpuLifeCycle(MonteCarloSimulation* sim, const std::vector<int>& devices);
    ~GpuLifeCycle();

    // Before setupSimulation(): when the spatial grid is an octree built by a DensityTreePolicy over geometric dust media the
    // engine knows, replaces the policy object by one whose constructTree() runs sk_engine_build_octree on `device`
    // (DensityTreePolicy.cpp:242-309).  Returns false (and changes nothing) otherwise.
    static bool installDeviceTreeConstruction(MonteCarloSimulation* sim, int device);

    // returns an empty string when the configured simulation lies on the accelerated path, or the reason why not
    std::string unsupportedReason() const;
    bool mediaShareOneMix() const;  // all medium components have the tables of one dust mix (they then run as one medium)

    // hands all tables to the engine (the extractor of INTEGRATION.md section 1); call after setupSimulation()
    void configure();

    // MonteCarloSimulation::runSimulation() with the life cycle on the GPU
    void runSimulation();

    const sk_counters_t& counters() const { return _counters; }

private:
    void check(int rc) const;
    void configureEngine(int device);
    void prepareNccl();
    void runSegmentOnAll(size_t Npp, int primary, int peel, int store);
    void allReduce(int which);
    void communicateRadiationField(int primary);
    void runPrimaryEmission();
    void runSecondaryEmission();
    void runSecondaryEmissionIterations();
    void runPrimaryEmissionIterations();
    void runMergedEmissionIterations();
    bool logDustConvergence(int iter, double& prevLabsseco);
    bool logLoopConvergence(bool converged, int iter, int minIters, int maxIters);
    bool updateDynamicMediumState();
    void uploadMediumState(sk_engine_t* engine);
    void returnRadiationField();
    void returnDetectors();

    MonteCarloSimulation* _sim;
    std::vector<int> _devices;
    std::vector<sk_engine_t*> _engines;  // one per device; _e is the one being configured, then the first
    std::vector<void*> _comms;           // ncclComm_t per engine (empty for a single device)
    sk_engine_t* _e{nullptr};
    uint32_t _segment{0};
    sk_counters_t _counters{};
};

#endif
