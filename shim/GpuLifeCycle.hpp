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
    GpuLifeCycle(MonteCarloSimulation* sim, int device);
    ~GpuLifeCycle();

    // returns an empty string when the configured simulation lies on the accelerated path, or the reason why not
    std::string unsupportedReason() const;

    // hands all tables to the engine (the extractor of INTEGRATION.md section 1); call after setupSimulation()
    void configure();

    // MonteCarloSimulation::runSimulation() with the life cycle on the GPU
    void runSimulation();

    const sk_counters_t& counters() const { return _counters; }

private:
    void check(int rc) const;
    void runPrimaryEmission();
    void runSecondaryEmission();
    void runSecondaryEmissionIterations();
    void returnRadiationField();
    void returnDetectors();

    MonteCarloSimulation* _sim;
    int _device;
    sk_engine_t* _e{nullptr};
    uint32_t _segment{0};
    sk_counters_t _counters{};
};

#endif
