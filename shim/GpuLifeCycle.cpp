// GpuLifeCycle.cpp -- see GpuLifeCycle.hpp.  C++14 like the reference.
//
// The data the engine needs lives in private members of a dozen reference classes.  A maintainer would
// add `friend class GpuLifeCycle;` to those classes; because this repository must not modify or copy the reference, this
// one translation unit instead sees the reference headers with `private`/`protected` mapped to `public` (the standard
// library headers are included first so that they are unaffected; access specifiers do not change object layout).

// ---- standard library first (see above)
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
#include <chrono>
#include <typeinfo>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <valarray>
#include <vector>

#include <cuda_runtime_api.h>
#include <nccl.h>

#define private public
#define protected public
#include "AllCellsLibrary.hpp"
#include "BlackBodySED.hpp"
#include "CartesianSpatialGrid.hpp"
#include "Configuration.hpp"
#include "CylindricalVectorField.hpp"
#include "DefaultWavelengthDistribution.hpp"
#include "DensityTreePolicy.hpp"
#include "DisjointWavelengthGrid.hpp"
#include "DistantInstrument.hpp"
#include "DustMix.hpp"
#include "DynamicStateOptions.hpp"
#include "DynamicStateRecipe.hpp"
#include "ExpDiskGeometry.hpp"
#include "FatalError.hpp"
#include "FluxRecorder.hpp"
#include "FrameInstrument.hpp"
#include "FullInstrument.hpp"
#include "GeometricMedium.hpp"
#include "GeometricSource.hpp"
#include "InstrumentSystem.hpp"
#include "Log.hpp"
#include "Medium.hpp"
#include "MediumSystem.hpp"
#include "MonteCarloSimulation.hpp"
#include "OctTreeNode.hpp"
#include "OligoWavelengthDistribution.hpp"
#include "PointSource.hpp"
#include "PolicyTreeSpatialGrid.hpp"
#include "ProbeSystem.hpp"
#include "ProcessManager.hpp"
#include "RadialVectorField.hpp"
#include "Random.hpp"
#include "RingGeometry.hpp"
#include "SEDInstrument.hpp"
#include "SecondarySourceSystem.hpp"
#include "ShellGeometry.hpp"
#include "SourceSystem.hpp"
#include "SpiralStructureGeometryDecorator.hpp"
#include "StringUtils.hpp"
#include "TabulatedSED.hpp"
#include "TimeLogger.hpp"
#include "TreeSpatialGrid.hpp"
#include "UnidirectionalVectorField.hpp"
#include "Units.hpp"
#include "VoronoiMeshSnapshot.hpp"
#include "VoronoiMeshSpatialGrid.hpp"
#undef private
#undef protected

#include "GpuLifeCycle.hpp"

////////////////////////////////////////////////////////////////////

// VoronoiMeshSnapshot::Cell is a private nested class that the reference defines inside VoronoiMeshSnapshot.cpp:85-194, so
// its header only forward-declares it.  The shim needs the site position and the neighbour list of every cell, for which
// the reference offers no accessor; this definition repeats the DATA LAYOUT of that class (base Box; Vec _r, _c; double
// _volume; vector<int> _neighbors; Array _properties) so that the two members can be read.  A maintainer integrating the
// engine would add two accessors to VoronoiMeshSnapshot instead.
class VoronoiMeshSnapshot::Cell : public Box
{
public:
    Vec _r;
    Vec _c;
    double _volume;
    vector<int> _neighbors;
    Array _properties;
};

namespace
{
    // detector array ids of the reference (anonymous enum in FluxRecorder.cpp:26-56)
    enum { RefTotal = 0, RefTransparent, RefPrimaryDirect, RefPrimaryScattered, RefSecondaryDirect, RefSecondaryScattered,
           RefSecondaryTransparent, RefPrimaryScatteredLevel = 28 };

    int refComponent(int c)
    {
        switch (c)
        {
            case SK_COMP_TOTAL: return RefTotal;
            case SK_COMP_TRANSPARENT: return RefTransparent;
            case SK_COMP_PRIMARY_DIRECT: return RefPrimaryDirect;
            case SK_COMP_PRIMARY_SCATTERED: return RefPrimaryScattered;
            case SK_COMP_SECONDARY_DIRECT: return RefSecondaryDirect;
            case SK_COMP_SECONDARY_SCATTERED: return RefSecondaryScattered;
            case SK_COMP_SECONDARY_TRANSPARENT: return RefSecondaryTransparent;
            default: return RefPrimaryScatteredLevel + (c - SK_COMP_PRIMARY_SCATTERED_LEVEL);
        }
    }

    const double* ptr(const Array& a) { return &a[0]; }

    bool sameValues(const Array& a, const Array& b)
    {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i != a.size(); ++i)
            if (a[i] != b[i]) return false;
        return true;
    }

    // the geometries that have a device-side sampler (include/sk_engine.h sk_geometry_kind)
    bool fillGeometry(const Geometry* geom, sk_source_t& s)
    {
        if (auto g = dynamic_cast<const ShellGeometry*>(geom))
        {
            s.geometry = SK_GEOM_SHELL;
            double p[] = {g->minRadius(), g->maxRadius(), g->exponent(), g->_smin, g->_sdiff, g->_tmin, g->_tmax};
            std::copy(p, p + 7, s.geom_params);
            return true;
        }
        if (auto g = dynamic_cast<const ExpDiskGeometry*>(geom))
        {
            s.geometry = SK_GEOM_EXPDISK;
            double p[] = {g->scaleLength(), g->scaleHeight(), g->minRadius(), g->maxRadius(), g->maxZ()};
            std::copy(p, p + 5, s.geom_params);
            return true;
        }
        if (auto g = dynamic_cast<const RingGeometry*>(geom))
        {
            s.geometry = SK_GEOM_RING;
            double p[] = {g->ringRadius(), g->width(), g->height()};
            std::copy(p, p + 3, s.geom_params);
            s.geom_table_n = static_cast<int>(g->_Rv.size());
            s.geom_table_x = ptr(g->_Rv);
            s.geom_table_P = ptr(g->_Xv);
            return true;
        }
        if (auto g = dynamic_cast<const SpiralStructureGeometryDecorator*>(geom))
        {
            auto d = dynamic_cast<const ExpDiskGeometry*>(g->geometry());
            if (!d) return false;
            s.geometry = SK_GEOM_SPIRAL_EXPDISK;
            double p[] = {d->scaleLength(), d->scaleHeight(), d->minRadius(), d->maxRadius(), d->maxZ(),
                          static_cast<double>(g->numArms()), g->_tanp, g->radiusZeroPoint(), g->phaseZeroPoint(),
                          g->perturbationWeight(), static_cast<double>(g->index()), g->_cn};
            std::copy(p, p + 12, s.geom_params);
            return true;
        }
        return false;
    }
}

////////////////////////////////////////////////////////////////////

// SURVEY.md 8f row f2 in the drop-in: DensityTreePolicy::constructTree (DensityTreePolicy.cpp:242-309) on the GPU.  The
// reference calls the policy through the virtual TreePolicy::constructTree, so a subclass that overrides just that function
// can stand in for the policy object the ski file created; everything else (properties, setupSelfBefore with the media
// lists, the dust mass and kappa) is the reference's own code, inherited.
namespace
{
    // a medium component as sk_engine_build_octree sees it: geometry kind + the parameters of Geometry::density
    bool describeMedium(const Medium* medium, sk_density_geometry_t& d, bool setupDone)
    {
        memset(&d, 0, sizeof d);
        auto gm = dynamic_cast<const GeometricMedium*>(medium);
        if (!gm) return false;
        const Geometry* geom = gm->geometry();
        if (setupDone)
        {
            d.number = gm->number();
            d.mass = gm->mass();
        }
        if (auto g = dynamic_cast<const ShellGeometry*>(geom))
        {
            d.geometry = SK_GEOM_SHELL;
            double p[] = {g->minRadius(), g->maxRadius(), g->exponent(), g->_A};
            std::copy(p, p + 4, d.p);
            return true;
        }
        if (auto g = dynamic_cast<const ExpDiskGeometry*>(geom))
        {
            d.geometry = SK_GEOM_EXPDISK;
            double p[] = {g->scaleLength(), g->scaleHeight(), g->minRadius(), g->maxRadius(), g->maxZ(), g->_rho0};
            std::copy(p, p + 6, d.p);
            return true;
        }
        if (auto g = dynamic_cast<const RingGeometry*>(geom))
        {
            d.geometry = SK_GEOM_RING;
            double p[] = {g->ringRadius(), g->width(), g->height(), g->_A};
            std::copy(p, p + 4, d.p);
            return true;
        }
        if (auto g = dynamic_cast<const SpiralStructureGeometryDecorator*>(geom))
        {
            auto e = dynamic_cast<const ExpDiskGeometry*>(g->geometry());
            if (!e) return false;
            d.geometry = SK_GEOM_SPIRAL_EXPDISK;
            double p[] = {e->scaleLength(), e->scaleHeight(), e->minRadius(), e->maxRadius(), e->maxZ(), e->_rho0,
                          static_cast<double>(g->numArms()), g->_tanp, g->radiusZeroPoint(), g->phaseZeroPoint(),
                          g->perturbationWeight(), static_cast<double>(g->index()), g->_cn};
            std::copy(p, p + 13, d.p);
            return true;
        }
        return false;
    }

    class GpuDensityTreePolicy : public DensityTreePolicy
    {
    public:
        GpuDensityTreePolicy(const DensityTreePolicy* src, int device) : _device(device)
        {
            _minLevel = src->_minLevel;
            _maxLevel = src->_maxLevel;
            _maxDustFraction = src->_maxDustFraction;
            _maxDustOpticalDepth = src->_maxDustOpticalDepth;
            _wavelength = src->_wavelength;
            _maxDustDensityDispersion = src->_maxDustDensityDispersion;
            _maxElectronFraction = src->_maxElectronFraction;
            _maxGasFraction = src->_maxGasFraction;
        }

        vector<TreeNode*> constructTree(TreeNode* root) override
        {
            auto log = find<Log>();
            // only dust criteria on geometric media run on the device; anything else is the reference's own loop
            vector<sk_density_geometry_t> media(_dustMedia.size());
            bool ok = !_dustMedia.empty() && _electronMedia.empty() && _gasMedia.empty() && _dustMIBv.empty()
                      && maxLevel() <= 15 && dynamic_cast<OctTreeNode*>(root);
            for (size_t h = 0; ok && h != media.size(); ++h) ok = describeMedium(_dustMedia[h], media[h], true);
            if (!ok) return DensityTreePolicy::constructTree(root);

            auto fail = [](int rc) {
                if (rc != SK_OK) throw FATALERROR(string("GPU tree construction: ") + sk_last_error());
            };
            auto now = []() { return std::chrono::steady_clock::now(); };
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
                return StringUtils::toString(std::chrono::duration<double, std::milli>(b - a).count(), 'f', 1);
            };
            auto t0 = now();
            sk_config_t c;
            memset(&c, 0, sizeof c);
            c.seed = static_cast<uint32_t>(_random->seed());
            c.device = _device;
            sk_engine_t* e = nullptr;
            fail(sk_engine_create(&c, &e));  // the first CUDA call of the process: creates the device context
            auto t1 = now();
            sk_tree_policy_t p;
            memset(&p, 0, sizeof p);
            p.min_level = minLevel();
            p.max_level = maxLevel();
            p.num_samples = _numSamples;
            p.max_dust_fraction = maxDustFraction();
            p.max_dust_optical_depth = maxDustOpticalDepth();
            p.max_dust_density_dispersion = maxDustDensityDispersion();
            p.dust_kappa = _dustKappa;
            const Box& b = *root;
            double ext[6] = {b.xmin(), b.ymin(), b.zmin(), b.xmax(), b.ymax(), b.zmax()};
            uint64_t nn = 0, nc = 0;
            int rc = sk_engine_build_octree(e, ext, &p, static_cast<int32_t>(media.size()), media.data(), &nn, &nc);
            vector<int32_t> firstChild(rc == SK_OK ? nn : 0);
            if (rc == SK_OK) rc = sk_engine_read_octree(e, firstChild.data());
            sk_engine_destroy(e);
            fail(rc);
            auto t2 = now();

            // the reference's node objects in the same breadth-first order (TreeNode::subdivide without the neighbour lists:
            // nothing on the host walks the tree from cell to cell once the life cycle runs on the device)
            vector<TreeNode*> nodev{root};
            nodev.reserve(nn);
            for (size_t l = 0; l != nn; ++l)
            {
                if (firstChild[l] < 0) continue;
                if (static_cast<size_t>(firstChild[l]) != nodev.size())
                    throw FATALERROR("GPU tree construction: node list is not breadth-first");
                TreeNode* node = nodev[l];
                node->createChildren(static_cast<int>(nodev.size()));
                nodev.insert(nodev.end(), node->_children.begin(), node->_children.end());
            }
            if (nodev.size() != nn) throw FATALERROR("GPU tree construction: node count mismatch");
            log->info("  GPU tree construction: " + std::to_string(nn) + " nodes, " + std::to_string(nc) + " cells (CUDA context "
                      + ms(t0, t1) + " ms, sk_engine_build_octree + read-back " + ms(t1, t2) + " ms, TreeNode objects "
                      + ms(t2, now()) + " ms)");
            return nodev;
        }

    private:
        int _device;
    };
}

bool GpuLifeCycle::installDeviceTreeConstruction(MonteCarloSimulation* sim, int device)
{
    auto ms = sim->mediumSystem();
    if (!ms) return false;
    auto grid = dynamic_cast<PolicyTreeSpatialGrid*>(ms->grid());
    if (!grid || grid->treeType() != PolicyTreeSpatialGrid::TreeType::OctTree) return false;
    auto policy = grid->policy();
    if (!policy || typeid(*policy) != typeid(DensityTreePolicy)) return false;
    for (auto medium : ms->media())
    {
        sk_density_geometry_t d;
        if (!describeMedium(medium, d, false)) return false;
    }
    grid->ii_set_policy(new GpuDensityTreePolicy(static_cast<DensityTreePolicy*>(policy), device));
    return true;
}

////////////////////////////////////////////////////////////////////

GpuLifeCycle::GpuLifeCycle(MonteCarloSimulation* sim, const std::vector<int>& devices) : _sim(sim), _devices(devices)
{
    if (_devices.empty()) _devices.push_back(0);
}

GpuLifeCycle::~GpuLifeCycle()
{
    for (void* c : _comms) ncclCommDestroy(static_cast<ncclComm_t>(c));
    for (sk_engine_t* e : _engines) sk_engine_destroy(e);
}

// every device gets a full replica of the model, like every process of the reference's MPI runs
void GpuLifeCycle::configure()
{
    for (int device : _devices)
    {
        _e = nullptr;
        configureEngine(device);
        _engines.push_back(_e);
    }
    _e = _engines[0];
    if (_engines.size() > 1)
    {
        for (size_t i = 0; i != _engines.size(); ++i)
            check(sk_engine_set_history_interleave(_engines[i], 16384, static_cast<uint32_t>(_engines.size()),
                                                   static_cast<uint32_t>(i)));
        prepareNccl();
    }
}

void GpuLifeCycle::prepareNccl()
{
    std::vector<ncclComm_t> comms(_devices.size());
    ncclResult_t r = ncclCommInitAll(comms.data(), static_cast<int>(_devices.size()), _devices.data());
    if (r != ncclSuccess) throw FATALERROR(string("NCCL: ") + ncclGetErrorString(r));
    for (ncclComm_t c : comms) _comms.push_back(c);
    _sim->log()->info("GPU life cycle: " + std::to_string(_devices.size()) + " devices, NCCL all-reduce of the tallies");
}

// one emission segment: contiguous history blocks, one per device, run concurrently
void GpuLifeCycle::runSegmentOnAll(size_t Npp, int primary, int peel, int store)
{
    const size_t n = _engines.size();
    const uint32_t segment = _segment++;
    if (n == 1)
    {
        check(sk_engine_run_segment(_e, 0, Npp, primary, peel, store, segment));
        return;
    }
    std::vector<int> rc(n, SK_OK);
    std::vector<string> msg(n);
    std::vector<std::thread> threads;
    for (size_t i = 0; i != n; ++i)
    {
        // every engine is given the whole segment and runs its interleaved share of it (set in configure():
        // sk_engine_set_history_interleave), which keeps the devices balanced also where the histories are ordered by cell
        threads.emplace_back([this, i, Npp, primary, peel, store, segment, &rc, &msg]() {
            rc[i] = sk_engine_run_segment(_engines[i], 0, Npp, primary, peel, store, segment);
            if (rc[i] != SK_OK) msg[i] = sk_last_error();  // the error text is thread-local
        });
    }
    for (auto& t : threads) t.join();
    for (size_t i = 0; i != n; ++i)
        if (rc[i] != SK_OK) throw FATALERROR("GPU life-cycle engine (device " + std::to_string(_devices[i]) + "): " + msg[i]);
}

// in-place sum over the devices of one tally block (sk_engine_device_buffer: 0 rf1, 2 rf2c, 3 detectors, 4 statistics)
void GpuLifeCycle::allReduce(int which)
{
    const size_t n = _engines.size();
    if (n == 1) return;
    std::vector<void*> buf(n), stream(n);
    uint64_t count = 0;
    for (size_t i = 0; i != n; ++i)
    {
        uint64_t c = 0;
        check(sk_engine_device_buffer(_engines[i], which, &buf[i], &c));
        check(sk_engine_cuda_stream(_engines[i], &stream[i]));
        if (i && c != count) throw FATALERROR("GPU life cycle: tally blocks of different size on different devices");
        count = c;
        if (!buf[i]) count = 0;
    }
    if (!count) return;
    ncclGroupStart();
    for (size_t i = 0; i != n; ++i)
    {
        ncclResult_t r = ncclAllReduce(buf[i], buf[i], count, ncclDouble, ncclSum, static_cast<ncclComm_t>(_comms[i]),
                                       static_cast<cudaStream_t>(stream[i]));
        if (r != ncclSuccess)
        {
            ncclGroupEnd();
            throw FATALERROR(string("NCCL all-reduce: ") + ncclGetErrorString(r));
        }
    }
    ncclResult_t r = ncclGroupEnd();
    if (r != ncclSuccess) throw FATALERROR(string("NCCL all-reduce: ") + ncclGetErrorString(r));
    for (size_t i = 0; i != n; ++i) check(sk_engine_synchronize(_engines[i]));
}

// MediumSystem::communicateRadiationField(primary), MediumSystem.cpp:1304-1313: sum over the devices, then _rf2 = _rf2c
void GpuLifeCycle::communicateRadiationField(int primary)
{
    allReduce(primary ? 0 : 2);
    for (sk_engine_t* e : _engines) check(sk_engine_communicate_rf(e, primary));
}

void GpuLifeCycle::check(int rc) const
{
    // the reference reports errors as FatalError exceptions (SkirtCommandLineHandler.cpp:372-400)
    if (rc != SK_OK) throw FATALERROR(string("GPU life-cycle engine: ") + sk_last_error());
}

////////////////////////////////////////////////////////////////////

// true when every medium component has the tables of the first one's dust mix
bool GpuLifeCycle::mediaShareOneMix() const
{
    auto ms = _sim->mediumSystem();
    auto mix = dynamic_cast<const DustMix*>(ms->media()[0]->mix());
    for (int h = 1; h < ms->numMedia(); ++h)
    {
        auto other = dynamic_cast<const DustMix*>(ms->media()[h]->mix());
        if (!mix || !other || other->type() != mix->type() || other->scatteringMode() != mix->scatteringMode()
            || other->mass() != mix->mass() || !sameValues(other->_lambdav, mix->_lambdav)
            || !sameValues(other->_sigmaabsv, mix->_sigmaabsv) || !sameValues(other->_sigmascav, mix->_sigmascav)
            || !sameValues(other->_asymmparv, mix->_asymmparv))
            return false;
    }
    return true;
}

namespace
{
    // GeometricSource::velocityMagnitude() * velocityDistribution()->vector(r) (GeometricSource.cpp:73-79) for the vector
    // fields the engine evaluates at the launch position
    bool fillVelocity(const GeometricSource* gs, sk_source_t& s)
    {
        auto field = gs->velocityDistribution();
        if (auto u = dynamic_cast<const UnidirectionalVectorField*>(field))
        {
            Vec d = u->vector(Position());
            s.velocity_kind = SK_VEL_CONSTANT;
            s.velocity[0] = gs->velocityMagnitude() * d.x();
            s.velocity[1] = gs->velocityMagnitude() * d.y();
            s.velocity[2] = gs->velocityMagnitude() * d.z();
            return true;
        }
        // (exact types only: a subclass may evaluate another field)
        if (field->type() == "RadialVectorField")
        {
            auto r = dynamic_cast<const RadialVectorField*>(field);
            s.velocity_kind = SK_VEL_RADIAL;
            s.velocity[0] = gs->velocityMagnitude();
            s.velocity[1] = r->unityRadius();
            s.velocity[2] = r->exponent();
            return true;
        }
        if (field->type() == "CylindricalVectorField")
        {
            auto c = dynamic_cast<const CylindricalVectorField*>(field);
            s.velocity_kind = SK_VEL_CYLINDRICAL;
            s.velocity[0] = gs->velocityMagnitude();
            s.velocity[1] = c->unityRadius();
            s.velocity[2] = c->exponent();
            return true;
        }
        return false;
    }
}

std::string GpuLifeCycle::unsupportedReason() const
{
    auto config = _sim->_config;
    auto ms = _sim->mediumSystem();
    if (!config->hasMedium() || !ms) return "no medium";
    // moving media (and moving sources) run on the engine's path with per-cell perceived wavelengths (sk_engine_set_velocities);
    // the other reasons for spatially variable cross sections do not
    if (config->hasVariableMedia()) return "spatially variable material mixes";
    if (config->hubbleExpansionRate()) return "Hubble flow";
    if (!config->hasMovingMedia() && !config->hasSingleConstantSectionMedium() && !config->hasMultipleConstantSectionMedia())
        return "variable cross sections";
    if (config->hasPolarization()) return "polarization";
    // dynamic medium state: the recipes run on the host between the segments (the reference's own code on the radiation field the
    // engine hands back) and the new densities go to the engine; media that keep a dynamic state of their own (gas) do not
    if (config->hasPrimaryDynamicStateMedia() || config->hasSecondaryDynamicStateMedia()) return "media with a dynamic state of their own";
    if (config->hasDynamicStateRecipes())
        for (auto recipe : ms->dynamicStateOptions()->recipes())
            if (recipe->type() != "ClearDensityRecipe") return "dynamic state recipe " + recipe->type();
    if (config->hasGasEmission()) return "gas emission";
    if (config->hasStochasticDustEmission()) return "stochastic dust emission";
    if (ProcessManager::isMultiProc()) return "MPI (use one engine per rank through the C ABI instead)";
    auto mix = dynamic_cast<const DustMix*>(ms->media()[0]->mix());
    if (!mix || mix->scatteringMode() != DustMix::ScatteringMode::HenyeyGreenstein) return "a material mix other than a Henyey-Greenstein dust mix";
    // several media (MediumSystem.cpp:874-885, 697-767): components that share one material mix are a single medium with the
    // summed density (opacity, albedo, phase function and emissivity are the same); components with different mixes run on the
    // engine's several-component path (sk_engine_set_media), up to SK_MAX_MEDIA of them
    for (int h = 1; h < ms->numMedia(); ++h)
    {
        auto other = dynamic_cast<const DustMix*>(ms->media()[h]->mix());
        if (!other || other->scatteringMode() != DustMix::ScatteringMode::HenyeyGreenstein)
            return "a material mix other than a Henyey-Greenstein dust mix";
        if (!sameValues(other->_lambdav, mix->_lambdav)) return "dust mixes with different property wavelength grids";
    }
    if (!mediaShareOneMix())
    {
        if (ms->numMedia() > SK_MAX_MEDIA) return "more than " + std::to_string(SK_MAX_MEDIA) + " media with different material mixes";
    }
    auto grid = ms->grid();
    auto tree = dynamic_cast<TreeSpatialGrid*>(grid);
    auto voronoi = dynamic_cast<VoronoiMeshSpatialGrid*>(grid);
    if (!dynamic_cast<CartesianSpatialGrid*>(grid) && !(tree && dynamic_cast<OctTreeNode*>(tree->_nodev[0])) && !voronoi)
        return "spatial grid " + grid->type();
    for (auto source : _sim->sourceSystem()->sources())
    {
        auto ns = dynamic_cast<NormalizedSource*>(source);
        if (!ns) return "source " + source->type();
        if (!dynamic_cast<BlackBodySED*>(ns->sed()) && !dynamic_cast<TabulatedSED*>(ns->sed())) return "SED " + ns->sed()->type();
        if (!ns->_oligochromatic && ns->_xi && !dynamic_cast<DefaultWavelengthDistribution*>(ns->_biasDistribution))
            return "wavelength bias distribution " + ns->_biasDistribution->type();
        if (auto ps = dynamic_cast<PointSource*>(source))
        {
            if (ps->angularDistribution() || ps->polarizationProfile()) return "anisotropic or polarized point source";
        }
        else if (auto gs = dynamic_cast<GeometricSource*>(source))
        {
            sk_source_t probe;
            memset(&probe, 0, sizeof probe);
            if (!fillGeometry(gs->geometry(), probe)) return "source geometry " + gs->geometry()->type();
            if (gs->hasVelocity() && !fillVelocity(gs, probe)) return "source velocity field " + gs->velocityDistribution()->type();
        }
        else
            return "source " + source->type();
    }
    for (auto ins : _sim->instrumentSystem()->instruments())
    {
        if (!dynamic_cast<SEDInstrument*>(ins) && !dynamic_cast<FullInstrument*>(ins)
            && !(dynamic_cast<FrameInstrument*>(ins) && ins->type() == "FrameInstrument"))
            return "instrument " + ins->type();
        if (ins->recordPolarization()) return "polarization recording";
        if (!dynamic_cast<const DisjointWavelengthGrid*>(ins->_recorder->_lambdagrid)) return "instrument wavelength grid that is not disjoint";
    }
    if (_sim->instrumentSystem()->instruments().size() > 8) return "more than 8 instruments";
    for (auto ins : _sim->instrumentSystem()->instruments())
    {
        if (ins->numScatteringLevels() > 8) return "more than 8 individually recorded scattering levels";
        auto fi = dynamic_cast<FrameInstrument*>(ins);
        if (fi && ins->recordStatistics()
            && static_cast<double>(fi->numPixelsX()) * fi->numPixelsY() * ins->_recorder->_lambdagrid->numBins() > 2147483647.)
            return "per-pixel statistics on a frame with more than 2^31 pixel bins";
    }
    // probes that are called back for every launched packet (LaunchedPacketsProbe, SourceSystem.cpp:101-113,
    // SecondarySourceSystem.cpp:130-142): the engine launches on the device and never calls them
    if (!_sim->sourceSystem()->_callbackv.empty()) return "a probe with a launch call-back (e.g. LaunchedPacketsProbe)";
    if (_sim->_secondarySourceSystem && !_sim->_secondarySourceSystem->_callbackv.empty())
        return "a probe with a secondary launch call-back (e.g. LaunchedPacketsProbe)";
    if (tree)
    {
        if (tree->_nodev.size() > 0x03FFFFFFu) return "octree with more than 2^26 nodes";
        int maxlevel = 0;
        for (auto node : tree->_nodev) maxlevel = std::max(maxlevel, node->level());
        if (maxlevel > 15) return "octree deeper than 15 levels";
    }
    if (config->hasSecondaryEmission())
    {
        if (!config->hasDustEmission()) return "secondary emission other than dust";
        if (!dynamic_cast<AllCellsLibrary*>(config->cellLibrary())) return "spatial cell library " + config->cellLibrary()->type();
        if (config->dustEmissionWavelengthBias()
            && !dynamic_cast<DefaultWavelengthDistribution*>(config->dustEmissionWavelengthBiasDistribution()))
            return "dust emission wavelength bias distribution";
        if (_sim->_secondarySourceSystem->numSources() != 1) return "more than one secondary source";
    }
    return std::string();
}

////////////////////////////////////////////////////////////////////

// The medium state the engine walks through: number densities per component and cell volumes (MediumState.cpp:196-247) and,
// with moving media, the bulk velocities.  Called when the engine is configured and again after every update of a dynamic
// medium state (MediumSystem::updateDynamicStateRecipes, MediumSystem.cpp:1498-1558), which changes densities on the host.
void GpuLifeCycle::uploadMediumState(sk_engine_t* engine)
{
    auto config = _sim->_config;
    auto ms = _sim->mediumSystem();
    sk_engine_t* _e = engine;
    int M = ms->numCells();
    const int numMedia = ms->numMedia();
    // components with the same material mix: the densities add up to one medium; different mixes: one engine component each
    const int numComponents = mediaShareOneMix() ? 1 : numMedia;
    vector<double> nv(static_cast<size_t>(numComponents) * M), Vv(M);
    for (int m = 0; m != M; ++m)
    {
        if (numComponents == 1)
        {
            nv[m] = ms->numberDensity(m, 0);
            for (int h = 1; h < numMedia; ++h) nv[m] += ms->numberDensity(m, h);
        }
        else
            for (int h = 0; h < numMedia; ++h) nv[static_cast<size_t>(h) * M + m] = ms->numberDensity(m, h);
        Vv[m] = ms->volume(m);
    }
    check(sk_engine_set_media(_e, M, numComponents, nv.data(), Vv.data()));
    if (config->hasMovingMedia())
    {
        // MediumState::bulkVelocity(m): the aggregate over the components the reference's set-up has stored per cell
        // (MediumSystem.cpp:330-365)
        vector<double> vv(3 * static_cast<size_t>(M));
        for (int m = 0; m != M; ++m)
        {
            Vec v = ms->bulkVelocity(m);
            vv[3 * static_cast<size_t>(m)] = v.x();
            vv[3 * static_cast<size_t>(m) + 1] = v.y();
            vv[3 * static_cast<size_t>(m) + 2] = v.z();
        }
        check(sk_engine_set_velocities(_e, M, vv.data()));
    }
}

void GpuLifeCycle::configureEngine(int device)
{
    auto config = _sim->_config;
    auto ms = _sim->mediumSystem();

    // ---- Configuration digest (Configuration.cpp:30-377)
    sk_config_t c;
    memset(&c, 0, sizeof c);
    c.seed = static_cast<uint32_t>(_sim->random()->seed());
    c.force_scattering = config->forceScattering();
    c.min_scatt_events = config->minScattEvents();
    c.path_length_bias = config->pathLengthBias();
    c.min_weight_reduction = config->minWeightReduction();
    c.explicit_absorption = config->explicitAbsorption();
    c.device = device;
    check(sk_engine_create(&c, &_e));

    // ---- spatial grid
    if (auto g = dynamic_cast<CartesianSpatialGrid*>(ms->grid()))
    {
        check(sk_engine_set_grid_cartesian(_e, g->_Nx, g->_Ny, g->_Nz, ptr(g->_xv), ptr(g->_yv), ptr(g->_zv)));
    }
    else if (auto v = dynamic_cast<VoronoiMeshSpatialGrid*>(ms->grid()))
    {
        // VoronoiMeshSnapshot::_cells: site positions and neighbour lists as built by the vendored voro++
        auto mesh = v->_mesh;
        size_t n = mesh->_cells.size();
        vector<double> sites(3 * n), boxes(6 * n);
        vector<int64_t> offset(n + 1, 0);
        vector<int32_t> index;
        // The shim reads the cells through its own declaration of the reference's private Cell class (top of this file).
        // Check that view against what the reference's PUBLIC accessors say about the same cells -- site position,
        // enclosing box, volume -- and the neighbour lists for plausibility (at least 4 faces, indices that exist): a
        // layout that has drifted apart ends the run here instead of feeding garbage to the engine.
        for (size_t m = 0; m < n; m += std::max<size_t>(1, n / 64))
        {
            const VoronoiMeshSnapshot::Cell* cell = mesh->_cells[m];
            const Position site = mesh->position(static_cast<int>(m));
            const Box box = mesh->extent(static_cast<int>(m));
            const Box& cb = *cell;
            bool same = cell->_r.x() == site.x() && cell->_r.y() == site.y() && cell->_r.z() == site.z()
                        && cell->_volume == mesh->volume(static_cast<int>(m)) && cb.xmin() == box.xmin()
                        && cb.ymin() == box.ymin() && cb.zmin() == box.zmin() && cb.xmax() == box.xmax()
                        && cb.ymax() == box.ymax() && cb.zmax() == box.zmax() && cell->_neighbors.size() >= 4
                        && cell->_neighbors.size() < 1000;
            for (size_t i = 0; same && i != cell->_neighbors.size(); ++i)
                same = cell->_neighbors[i] >= -6 && cell->_neighbors[i] < static_cast<int>(n);
            if (!same)
                throw FATALERROR("GPU life-cycle shim: the layout of VoronoiMeshSnapshot::Cell differs from the shim's "
                                 "declaration (cell " + std::to_string(m) + ")");
        }
        for (size_t m = 0; m != n; ++m)
        {
            const VoronoiMeshSnapshot::Cell* cell = mesh->_cells[m];
            sites[3 * m] = cell->_r.x();
            sites[3 * m + 1] = cell->_r.y();
            sites[3 * m + 2] = cell->_r.z();
            // the cell's enclosing box (its Box base class, VoronoiMeshSnapshot.cpp:104-135): dust emission positions
            const Box& cb = *cell;
            double cbox[6] = {cb.xmin(), cb.ymin(), cb.zmin(), cb.xmax(), cb.ymax(), cb.zmax()};
            std::copy(cbox, cbox + 6, boxes.begin() + 6 * m);
            index.insert(index.end(), cell->_neighbors.begin(), cell->_neighbors.end());
            offset[m + 1] = static_cast<int64_t>(index.size());
        }
        Box b = mesh->_extent;
        double ext[6] = {b.xmin(), b.ymin(), b.zmin(), b.xmax(), b.ymax(), b.zmax()};
        check(sk_engine_set_grid_voronoi(_e, ext, static_cast<int32_t>(n), sites.data(), offset.data(), index.data()));
        if (config->hasSecondaryEmission()) check(sk_engine_set_voronoi_extents(_e, static_cast<int32_t>(n), boxes.data()));
    }
    else
    {
        auto t = dynamic_cast<TreeSpatialGrid*>(ms->grid());
        size_t n = t->_nodev.size();
        vector<int32_t> firstChild(n);
        for (size_t l = 0; l != n; ++l)
        {
            auto node = t->_nodev[l];
            firstChild[l] = node->isChildless() ? -1 : node->children()[0]->id();
        }
        Box b = t->extent();
        double ext[6] = {b.xmin(), b.ymin(), b.zmin(), b.xmax(), b.ymax(), b.zmax()};
        check(sk_engine_set_grid_octree(_e, ext, static_cast<int32_t>(n), firstChild.data()));
    }

    // ---- medium state (MediumState.cpp:196-247) and, with moving media, the bulk velocities
    uploadMediumState(_e);
    const int numMedia = ms->numMedia();
    const int numComponents = mediaShareOneMix() ? 1 : numMedia;

    // ---- dust mix tables (DustMix.cpp:47-246), one set per engine component
    auto mix = dynamic_cast<const DustMix*>(ms->media()[0]->mix());
    vector<sk_dustmix_t> dv(numComponents);
    for (int h = 0; h < numComponents; ++h)
    {
        auto mh = dynamic_cast<const DustMix*>(ms->media()[h]->mix());
        sk_dustmix_t& d = dv[h];
        memset(&d, 0, sizeof d);
        d.num_lambda = static_cast<int32_t>(mh->_lambdav.size());
        d.lambda_border = ptr(mh->_lambdav);
        d.sigma_abs = ptr(mh->_sigmaabsv);
        d.sigma_sca = ptr(mh->_sigmascav);
        d.asymmpar = ptr(mh->_asymmparv);
        d.mu = mh->mass();
    }
    check(sk_engine_set_dustmixes(_e, numComponents, dv.data()));

    // ---- wavelength grids: every distinct grid used by the instruments, the radiation field and dust emission
    vector<DisjointWavelengthGrid*> grids;
    auto gridIndex = [&grids](DisjointWavelengthGrid* g) {
        auto it = std::find(grids.begin(), grids.end(), g);
        if (it != grids.end()) return static_cast<int>(it - grids.begin());
        grids.push_back(g);
        return static_cast<int>(grids.size()) - 1;
    };
    int rfGrid = config->hasRadiationField() ? gridIndex(config->radiationFieldWLG()) : -1;
    int emGrid = config->hasDustEmission() ? gridIndex(config->dustEmissionWLG()) : -1;
    vector<int> insGrid;
    for (auto ins : _sim->instrumentSystem()->instruments())
        insGrid.push_back(gridIndex(const_cast<DisjointWavelengthGrid*>(dynamic_cast<const DisjointWavelengthGrid*>(ins->_recorder->_lambdagrid))));
    vector<sk_wavelength_grid_t> wg(grids.size());
    vector<vector<int32_t>> ellv(grids.size());
    for (size_t i = 0; i != grids.size(); ++i)
    {
        auto g = grids[i];
        ellv[i].assign(g->_ellv.begin(), g->_ellv.end());
        wg[i].num_bins = g->numBins();
        wg[i].num_borders = static_cast<int32_t>(g->_borderv.size());
        wg[i].borders = ptr(g->_borderv);
        wg[i].ell = ellv[i].data();
        wg[i].lambda = ptr(g->_lambdav);
        wg[i].dlambda = ptr(g->_dlambdav);
    }
    check(sk_engine_set_wavelength_grids(_e, static_cast<int32_t>(wg.size()), wg.data(), rfGrid));

    // ---- primary sources (SourceSystem.cpp:14-41, NormalizedSource.cpp:20-70)
    auto ss = _sim->sourceSystem();
    vector<sk_source_t> sv(ss->sources().size());
    vector<double> oligoLambda;
    if (config->oligochromatic())
    {
        auto od = dynamic_cast<OligoWavelengthDistribution*>(config->oligoWavelengthBiasDistribution());
        oligoLambda.assign(begin(od->_wavelengths), end(od->_wavelengths));
    }
    for (size_t h = 0; h != sv.size(); ++h)
    {
        auto ns = dynamic_cast<NormalizedSource*>(ss->sources()[h]);
        sk_source_t& s = sv[h];
        memset(&s, 0, sizeof s);
        s.luminosity = ns->luminosity();
        s.source_weight = ns->sourceWeight();
        if (auto ps = dynamic_cast<PointSource*>(ns))
        {
            s.kind = SK_SRC_POINT;
            s.position[0] = ps->positionX();
            s.position[1] = ps->positionY();
            s.position[2] = ps->positionZ();
            if (ps->hasVelocity())  // SpecialtySource: velocityX/Y/Z
            {
                s.velocity_kind = SK_VEL_CONSTANT;
                s.velocity[0] = ps->velocityX();
                s.velocity[1] = ps->velocityY();
                s.velocity[2] = ps->velocityZ();
            }
        }
        else
        {
            s.kind = SK_SRC_GEOMETRIC;
            auto gs = dynamic_cast<GeometricSource*>(ns);
            fillGeometry(gs->geometry(), s);
            if (gs->hasVelocity()) fillVelocity(gs, s);
        }
        if (auto bb = dynamic_cast<BlackBodySED*>(ns->sed()))
        {
            s.sed_kind = SK_SED_BLACKBODY;
            s.sed_n = static_cast<int32_t>(bb->_lambdav.size());
            s.sed_lambda = ptr(bb->_lambdav);
            s.sed_p = ptr(bb->_pv);
            s.sed_P = ptr(bb->_Pv);
            s.sed_temperature = bb->temperature();
            s.sed_norm = bb->_Ltot;
        }
        else
        {
            auto ts = dynamic_cast<TabulatedSED*>(ns->sed());
            s.sed_kind = SK_SED_TABULATED;
            s.sed_n = static_cast<int32_t>(ts->_lambdav.size());
            s.sed_lambda = ptr(ts->_lambdav);
            s.sed_p = ptr(ts->_pv);
            s.sed_P = ptr(ts->_Pv);
        }
        s.wavelength_bias = ns->_xi;
        if (ns->_oligochromatic)
        {
            auto od = dynamic_cast<OligoWavelengthDistribution*>(ns->_biasDistribution);
            s.bias_kind = SK_BIAS_OLIGO;
            s.oligo_n = static_cast<int32_t>(oligoLambda.size());
            s.oligo_lambda = oligoLambda.data();
            s.oligo_probability = od->_probability;
        }
        else if (ns->_xi)
        {
            auto dd = dynamic_cast<DefaultWavelengthDistribution*>(ns->_biasDistribution);
            s.bias_kind = SK_BIAS_LOGUNIFORM;
            s.bias_min = dd->_range.min();
            s.bias_max = dd->_range.max();
        }
    }
    check(sk_engine_set_sources(_e, static_cast<int32_t>(sv.size()), sv.data(), ss->sourceBias()));

    // ---- instruments (DistantInstrument.cpp:13-77, FrameInstrument.cpp:12-32, FluxRecorder.cpp:185-300)
    vector<sk_instrument_t> iv(_sim->instrumentSystem()->instruments().size());
    for (size_t i = 0; i != iv.size(); ++i)
    {
        auto ins = _sim->instrumentSystem()->instruments()[i];
        auto di = dynamic_cast<DistantInstrument*>(ins);
        sk_instrument_t& q = iv[i];
        memset(&q, 0, sizeof q);
        q.wavelength_grid = insGrid[i];
        q.inclination = di->inclination();
        q.azimuth = di->azimuth();
        q.roll = di->roll();
        q.distance = di->distance();
        if (auto si = dynamic_cast<SEDInstrument*>(ins))
        {
            q.kind = SK_INSTR_SED;
            q.radius = si->radius();
        }
        else
        {
            auto fi = dynamic_cast<FrameInstrument*>(ins);
            q.kind = dynamic_cast<FullInstrument*>(ins) ? SK_INSTR_FULL : SK_INSTR_FRAME;
            q.num_pixels_x = fi->numPixelsX();
            q.num_pixels_y = fi->numPixelsY();
            q.field_of_view_x = fi->fieldOfViewX();
            q.field_of_view_y = fi->fieldOfViewY();
            q.center_x = fi->centerX();
            q.center_y = fi->centerY();
        }
        q.record_components = !ins->_recorder->_recordTotalOnly;
        q.num_scattering_levels = ins->numScatteringLevels();
        q.record_statistics = ins->recordStatistics();
        q.redshift = ins->_recorder->_redshift;  // observer frame: packets are binned at lambda (1 + z), FluxRecorder.cpp:310
    }
    check(sk_engine_set_instruments(_e, static_cast<int32_t>(iv.size()), iv.data(), config->hasSecondaryEmission()));

    // ---- dust emission (EquilibriumDustEmissionCalculator.cpp:18-93)
    if (config->hasSecondaryEmission())
    {
        vector<sk_secondary_t> secv(numComponents);
        Range r = config->dustEmissionWLG()->wavelengthRange();
        for (int h = 0; h < numComponents; ++h)
        {
            auto& calc = dynamic_cast<const DustMix*>(ms->media()[h]->mix())->_calc;
            sk_secondary_t& sec = secv[h];
            memset(&sec, 0, sizeof sec);
            sec.emission_grid = emGrid;
            sec.num_temperatures = static_cast<int32_t>(calc._Tv.size());
            sec.spatial_bias = config->secondarySpatialBias();
            sec.wavelength_bias = config->dustEmissionWavelengthBias();
            sec.bias_min = r.min();
            sec.bias_max = r.max();
            sec.temperature = ptr(calc._Tv);
            sec.planck_abs = ptr(calc._planckabsvv[0]);
            sec.rf_sigma_abs = ptr(calc._rfsigmaabsvv[0]);
            sec.em_sigma_abs = ptr(calc._emsigmaabsvv[0]);
            // the CMB source term of the energy balance (all zero unless DustEmissionOptions::includeHeatingByCMB)
            sec.rf_cmb = calc._Bcmbv.size() ? ptr(calc._Bcmbv) : nullptr;
        }
        check(sk_engine_set_secondary_media(_e, numComponents, secv.data()));
    }
}

////////////////////////////////////////////////////////////////////

// MonteCarloSimulation::runSimulation, MonteCarloSimulation.cpp:58-100
void GpuLifeCycle::runSimulation()
{
    auto config = _sim->_config;
    {
        TimeLogger logger(_sim->log(), "the run");
        // MonteCarloSimulation::runSimulation, MonteCarloSimulation.cpp:58-86
        bool hasPrimaryLuminosity = _sim->sourceSystem()->luminosity() > 0.;
        if (config->hasMergedIterations() && hasPrimaryLuminosity)
        {
            if (config->hasPrimaryIterations()) runPrimaryEmissionIterations();
            runMergedEmissionIterations();
            runPrimaryEmission();
            runSecondaryEmission();
        }
        else
        {
            if (config->hasPrimaryIterations() && hasPrimaryLuminosity) runPrimaryEmissionIterations();
            runPrimaryEmission();
            if (config->hasSecondaryEmission())
            {
                if (config->hasSecondaryIterations()) runSecondaryEmissionIterations();
                runSecondaryEmission();
            }
        }
    }
    memset(&_counters, 0, sizeof _counters);
    for (sk_engine_t* e : _engines)
    {
        sk_counters_t c;
        check(sk_engine_counters(e, &c, 0));
        const uint64_t* src = reinterpret_cast<const uint64_t*>(&c);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&_counters);
        for (size_t k = 0; k != sizeof(sk_counters_t) / sizeof(uint64_t); ++k) dst[k] += src[k];
    }
    // FluxRecorder::calibrateAndWrite -> ProcessManager::sumToRoot, FluxRecorder.cpp:487-493
    allReduce(3);
    allReduce(4);
    {
        TimeLogger logger(_sim->log(), "final output");
        returnRadiationField();
        _sim->probeSystem()->probeRun();
        returnDetectors();
        _sim->instrumentSystem()->flush();
        _sim->instrumentSystem()->write();
    }
}

// MonteCarloSimulation::runPrimaryEmission, MonteCarloSimulation.cpp:104-138
void GpuLifeCycle::runPrimaryEmission()
{
    auto config = _sim->_config;
    string segment = "primary emission";
    TimeLogger logger(_sim->log(), segment);
    if (config->hasRadiationField())
        for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 1));
    size_t Npp = config->numPrimaryPackets();
    if (!Npp)
        _sim->log()->warning("Skipping primary emission because no photon packets were requested");
    else if (!_sim->sourceSystem()->luminosity())
        _sim->log()->warning("Skipping primary emission because the total luminosity of primary sources is zero");
    else
    {
        _sim->log()->info("Launching " + StringUtils::toString(static_cast<double>(Npp)) + " primary emission photon packets on the GPU");
        for (sk_engine_t* e : _engines) check(sk_engine_prepare_primary(e, Npp));
        runSegmentOnAll(Npp, 1, 1, config->hasRadiationField());
    }
    if (config->hasRadiationField()) communicateRadiationField(1);
}

// MonteCarloSimulation::runSecondaryEmission, MonteCarloSimulation.cpp:142-173
void GpuLifeCycle::runSecondaryEmission()
{
    auto config = _sim->_config;
    string segment = "secondary emission";
    TimeLogger logger(_sim->log(), segment);
    bool storeRF = config->storeEmissionRadiationField();
    if (storeRF)
        for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 0));
    size_t Npp = config->numSecondaryPackets();
    double L = 0.;
    if (!Npp)
        _sim->log()->warning("Skipping secondary emission because no photon packets were requested");
    else
    {
        // every device prepares the same secondary sources from its (all-reduced) copy of the radiation field
        for (sk_engine_t* e : _engines) check(sk_engine_prepare_secondary(e, Npp, &L));
        if (!L)
            _sim->log()->warning("Skipping secondary emission because the total luminosity of secondary sources is zero");
        else
        {
            auto units = _sim->units();
            _sim->log()->info("Dust luminosity: " + StringUtils::toString(units->obolluminosity(L), 'g') + " " + units->ubolluminosity());
            runSegmentOnAll(Npp, 0, 1, storeRF);
        }
    }
    if (storeRF) communicateRadiationField(0);
}

// DustAbsorptionConvergence::logConvergenceInfo (MonteCarloSimulation.cpp:180-227): the absorbed luminosities from the
// engine's radiation field tables; the log lines are the reference's
bool GpuLifeCycle::logDustConvergence(int iter, double& prevLabsseco)
{
    auto config = _sim->_config;
    auto log = _sim->log();
    auto units = _sim->units();
    double fractionOfPrimary = config->maxFractionOfPrimary();
    double fractionOfPrevious = config->maxFractionOfPrevious();
    double Labsprim = 0., Labsseco = 0.;
    check(sk_engine_absorbed_luminosity(_e, 1, &Labsprim));
    check(sk_engine_absorbed_luminosity(_e, 0, &Labsseco));
    log->info("The total dust-absorbed primary luminosity is " + StringUtils::toString(units->obolluminosity(Labsprim), 'g')
              + " " + units->ubolluminosity());
    log->info("The total dust-absorbed secondary luminosity in iteration " + std::to_string(iter) + " is "
              + StringUtils::toString(units->obolluminosity(Labsseco), 'g') + " " + units->ubolluminosity());
    if (Labsprim > 0. && Labsseco > 0.)
    {
        if (iter == 1)
            log->info("--> absorbed secondary luminosity is " + StringUtils::toString(Labsseco / Labsprim * 100., 'f', 2)
                      + "% of absorbed primary luminosity (convergence criterion is "
                      + StringUtils::toString(fractionOfPrimary * 100., 'f', 2) + "%)");
        else
            log->info("--> absorbed secondary luminosity changed by "
                      + StringUtils::toString(std::abs((Labsseco - prevLabsseco) / Labsseco) * 100., 'f', 2)
                      + "% compared to previous iteration (convergence criterion is "
                      + StringUtils::toString(fractionOfPrevious * 100., 'f', 2) + "%)");
    }
    bool converged = Labsprim <= 0. || Labsseco <= 0. || Labsseco / Labsprim < fractionOfPrimary
                     || std::abs((Labsseco - prevLabsseco) / Labsseco) < fractionOfPrevious;
    prevLabsseco = Labsseco;
    return converged;
}

// logLoopConvergence (MonteCarloSimulation.cpp:233-261): true when the loop ends
bool GpuLifeCycle::logLoopConvergence(bool converged, int iter, int minIters, int maxIters)
{
    auto log = _sim->log();
    if (converged && iter < minIters)
    {
        log->info("Convergence reached but continuing until " + std::to_string(minIters) + " iterations have been performed");
        return false;
    }
    if (converged)
    {
        log->info("Convergence reached after " + std::to_string(iter) + " iterations");
        return true;
    }
    if (iter < maxIters)
    {
        log->info("Convergence not yet reached after " + std::to_string(iter) + " iterations");
        return false;
    }
    log->error("Convergence not yet reached after " + std::to_string(iter) + " iterations");
    return true;
}

// MediumSystem::updatePrimaryDynamicMediumState (MediumSystem.cpp:1626-1632) between two segments: the reference's recipes
// read the mean intensity from the reference's tables, so the engine's radiation field goes back to the host first; the
// densities they change go to every engine afterwards
bool GpuLifeCycle::updateDynamicMediumState()
{
    returnRadiationField();
    bool converged = _sim->mediumSystem()->updatePrimaryDynamicMediumState();
    for (sk_engine_t* e : _engines) uploadMediumState(e);
    return converged;
}

// MonteCarloSimulation::runPrimaryEmissionIterations, MonteCarloSimulation.cpp:266-330
void GpuLifeCycle::runPrimaryEmissionIterations()
{
    auto config = _sim->_config;
    auto log = _sim->log();
    double minNpp = std::max(1., config->numPrimaryIterationPackets() * config->primaryIterationInitialPacketsFraction());
    double maxNpp = std::max(1., static_cast<double>(config->numPrimaryIterationPackets()));
    double ramp = config->primaryIterationPacketsRamp();
    int minIters = config->minPrimaryIterations();
    int maxIters = config->maxPrimaryIterations();
    size_t prevNpp = 0;
    int iter = 0;
    while (true)
    {
        ++iter;
        size_t Npp = std::min(maxNpp, minNpp * std::pow(ramp, iter - 1));
        if (Npp != prevNpp)
        {
            for (sk_engine_t* e : _engines) check(sk_engine_prepare_primary(e, Npp));
            prevNpp = Npp;
        }
        bool converged = true;
        {
            string segment = "primary emission iteration " + std::to_string(iter);
            TimeLogger logger(log, segment);
            _sim->mediumSystem()->beginDynamicMediumStateIteration();
            for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 1));
            log->info("Launching " + StringUtils::toString(static_cast<double>(Npp)) + " primary emission photon packets on the GPU");
            runSegmentOnAll(Npp, 1, 0, 1);
            communicateRadiationField(1);
            converged = updateDynamicMediumState();
        }
        _sim->probeSystem()->probePrimary(iter);
        if (logLoopConvergence(converged, iter, minIters, maxIters)) break;
    }
}

// MonteCarloSimulation::runMergedEmissionIterations, MonteCarloSimulation.cpp:407-496
void GpuLifeCycle::runMergedEmissionIterations()
{
    auto config = _sim->_config;
    auto log = _sim->log();
    auto units = _sim->units();
    size_t Npp1 = config->numPrimaryIterationPackets();
    size_t Npp2 = config->numSecondaryIterationPackets();
    int minIters = config->minSecondaryIterations();
    int maxIters = config->maxSecondaryIterations();
    for (sk_engine_t* e : _engines) check(sk_engine_prepare_primary(e, Npp1));
    double prevLabsseco = 0.;
    int iter = 0;
    while (true)
    {
        ++iter;
        bool converged = true;
        {
            string segment = "merged primary and secondary emission iteration " + std::to_string(iter);
            TimeLogger logger(log, segment);
            _sim->mediumSystem()->beginDynamicMediumStateIteration();
            for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 1));
            runSegmentOnAll(Npp1, 1, 0, 1);
            communicateRadiationField(1);
            // (updateSecondaryDynamicMediumState, MediumSystem.cpp:1636-1641: only media with a dynamic state of their own,
            //  which unsupportedReason() excludes)
            for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 0));
            double L = 0.;
            for (sk_engine_t* e : _engines) check(sk_engine_prepare_secondary(e, Npp2, &L));
            if (!L)
            {
                log->warning("Skipping merged emission iterations because the total luminosity of secondary sources is zero");
                return;
            }
            log->info("Dust luminosity: " + StringUtils::toString(units->obolluminosity(L), 'g') + " " + units->ubolluminosity());
            runSegmentOnAll(Npp2, 0, 0, 1);
            communicateRadiationField(0);
            // (in the reference's order: the absorbed luminosities are evaluated with the updated densities)
            converged &= updateDynamicMediumState();
            converged &= logDustConvergence(iter, prevLabsseco);
        }
        _sim->probeSystem()->probeSecondary(iter);
        if (logLoopConvergence(converged, iter, minIters, maxIters)) break;
    }
}

// MonteCarloSimulation::runSecondaryEmissionIterations (.cpp:335-403)
void GpuLifeCycle::runSecondaryEmissionIterations()
{
    auto config = _sim->_config;
    auto log = _sim->log();
    auto units = _sim->units();
    size_t Npp = config->numSecondaryIterationPackets();
    int minIters = config->minSecondaryIterations();
    int maxIters = config->maxSecondaryIterations();
    double prevLabsseco = 0.;
    int iter = 0;
    while (true)
    {
        ++iter;
        bool converged = true;
        {
            string segment = "secondary emission iteration " + std::to_string(iter);
            TimeLogger logger(log, segment);
            for (sk_engine_t* e : _engines) check(sk_engine_clear_rf(e, 0));
            double L = 0.;
            for (sk_engine_t* e : _engines) check(sk_engine_prepare_secondary(e, Npp, &L));
            if (!L)
            {
                log->warning("Skipping secondary emission iterations because the total luminosity of secondary sources is zero");
                return;
            }
            log->info("Dust luminosity: " + StringUtils::toString(units->obolluminosity(L), 'g') + " " + units->ubolluminosity());
            runSegmentOnAll(Npp, 0, 0, 1);
            communicateRadiationField(0);
            converged = logDustConvergence(iter, prevLabsseco);
        }
        // probes that fire after every iteration read the radiation field from the reference's tables
        returnRadiationField();
        _sim->probeSystem()->probeSecondary(iter);
        if (logLoopConvergence(converged, iter, minIters, maxIters)) break;
    }
}

////////////////////////////////////////////////////////////////////

// the radiation field tables of the reference (MediumSystem.hpp:883-885) for its probes
void GpuLifeCycle::returnRadiationField()
{
    auto ms = _sim->mediumSystem();
    if (!_sim->_config->hasRadiationField()) return;
    if (ms->_rf1.size()) check(sk_engine_read_rf(_e, 0, &ms->_rf1.data()[0]));
    if (ms->_rf2.size()) check(sk_engine_read_rf(_e, 1, &ms->_rf2.data()[0]));
}

// the detector arrays of every FluxRecorder (FluxRecorder.hpp:396-404) before FluxRecorder::calibrateAndWrite
void GpuLifeCycle::returnDetectors()
{
    auto& instruments = _sim->instrumentSystem()->instruments();
    for (size_t i = 0; i != instruments.size(); ++i)
    {
        auto rec = instruments[i]->_recorder;
        int numComp = static_cast<int>(rec->_sed.size());
        for (int c = 0; c != SK_COMP_PRIMARY_SCATTERED_LEVEL + 8; ++c)
        {
            int rc = refComponent(c);
            if (rc >= numComp) continue;
            // in the reference the Total array exists only when components are not recorded (FluxRecorder.cpp:214-218)
            if (rec->_sed[rc].size()) check(sk_engine_read_sed(_e, static_cast<int32_t>(i), c, &rec->_sed[rc][0]));
            if (rec->_ifu[rc].size()) check(sk_engine_read_ifu(_e, static_cast<int32_t>(i), c, &rec->_ifu[rc][0]));
        }
        for (size_t k = 0; k != rec->_wsed.size(); ++k)
            if (rec->_wsed[k].size()) check(sk_engine_read_sed_stats(_e, static_cast<int32_t>(i), static_cast<int32_t>(k), &rec->_wsed[k][0]));
        for (size_t k = 0; k != rec->_wifu.size(); ++k)
            if (rec->_wifu[k].size()) check(sk_engine_read_ifu_stats(_e, static_cast<int32_t>(i), static_cast<int32_t>(k), &rec->_wifu[k][0]));
    }
}
