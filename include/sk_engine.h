/* sk_engine.h -- C ABI of the B200 photon-packet life-cycle engine (libskirt9_b200.so).
 *
 * This is the drop-in boundary for the one hot path of SKIRT 9 that this repository accelerates:
 * MonteCarloSimulation::performLifeCycle() and everything it calls.  The
 * reference has no C ABI of its own; the seam is the six call sites
 *     parallel->call(Npp, [this](size_t i, size_t n){ performLifeCycle(i, n, primary, peel, store); })
 * in SKIRT/core/MonteCarloSimulation.cpp:126-128,165-166,314,381,451,474, each followed by
 * instrumentSystem()->flush(), wait() and mediumSystem()->communicateRadiationField().  Every entry
 * point below names the reference function(s) whose work it takes over (paths relative to the
 * reference root).  INTEGRATION.md shows the C++ shim a SKIRT maintainer would add on their side.
 *
 * Conventions: plain C types only; the caller owns every input array (the engine copies it to the
 * device during the call); outputs are copied into caller-allocated host buffers; every function
 * returns SK_OK or an error code and sk_last_error() gives the message (the reference throws
 * FatalError; the shim rethrows).  All quantities are SI (m, W, rad), exactly the internal units of
 * the reference.  All entry points are to be called from one host thread (the reference calls MPI
 * from the parent thread only, SKIRT/mpi/ProcessManager.cpp:45).  There is NO CPU fallback: if no
 * CUDA device is present sk_engine_create() fails with SK_ERR_CUDA.
 *
 * The same structs and the same function set (prefix sko_ instead of sk_engine_) are implemented by
 * the CPU oracle in oracle/sk_oracle.c, which is test infrastructure only.
 */
#ifndef SK_ENGINE_H
#define SK_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_ABI_VERSION 6
#define SK_MAX_MEDIA 4 /* medium components with their own material mix (sk_engine_set_media) */

typedef struct sk_engine sk_engine_t;

enum sk_status {
    SK_OK = 0,
    SK_ERR_INVALID = 1,     /* bad argument (reference: FATALERROR in setup) */
    SK_ERR_UNSUPPORTED = 2, /* configuration outside the accelerated path (SURVEY.md 8a "OUT OF SCOPE") */
    SK_ERR_CUDA = 3,        /* CUDA runtime failure / no device */
    SK_ERR_STATE = 4        /* call sequence violated (e.g. run before grid/medium/sources are set) */
};

/* ---- Configuration: the digest of SKIRT/core/Configuration.cpp:30-377 the hot path reads ------- */
typedef struct sk_config {
    uint32_t seed;               /* Random::seed() (SKIRT/core/Random.hpp); Philox key word 0 */
    int32_t force_scattering;    /* Configuration::forceScattering() (PhotonPacketOptions.hpp:63) */
    int32_t min_scatt_events;    /* Configuration::minScattEvents() (PhotonPacketOptions.hpp:75) */
    double path_length_bias;     /* Configuration::pathLengthBias() (PhotonPacketOptions.hpp:83) */
    double min_weight_reduction; /* Configuration::minWeightReduction() (PhotonPacketOptions.hpp:69) */
    int32_t device;              /* CUDA device ordinal this engine instance owns (one engine per GPU) */
    int32_t explicit_absorption; /* Configuration::explicitAbsorption() (PhotonPacketOptions): interaction points are drawn
                                    in scattering optical depth and the packet is attenuated by exp(-tau_abs) instead of
                                    being multiplied with the albedo (MonteCarloSimulation.cpp:567-570, 727-731, 757-762) */
} sk_config_t;

/* ---- Wavelength grids: DisjointWavelengthGrid::bin (DisjointWavelengthGrid.cpp:332-341) -------- */
typedef struct sk_wavelength_grid {
    int32_t num_bins;       /* N = numBins() */
    int32_t num_borders;    /* K = size of _borderv (N+1, or 2N for disjoint bins) */
    const double* borders;  /* [K] ascending, DisjointWavelengthGrid::_borderv */
    const int32_t* ell;     /* [K+1] DisjointWavelengthGrid::_ellv : border index -> bin index or -1 */
    const double* lambda;   /* [N] characteristic wavelengths, _lambdav */
    const double* dlambda;  /* [N] effective widths, _dlambdav */
} sk_wavelength_grid_t;

/* ---- Dust mix: DustMix tables (DustMix.cpp:47-246) for the Henyey-Greenstein scattering mode ---- */
typedef struct sk_dustmix {
    int32_t num_lambda;          /* size of the property grid */
    int32_t reserved;
    const double* lambda_border; /* [n] DustMix::_lambdav (geometric-mean borders, DustMix.cpp:91-98) */
    const double* sigma_abs;     /* [n] DustMix::_sigmaabsv  (m2 per entity) */
    const double* sigma_sca;     /* [n] DustMix::_sigmascav */
    const double* asymmpar;      /* [n] DustMix::_asymmparv (already clamped to +-0.999999, DustMix.cpp:139-146) */
    double mu;                   /* MaterialMix::mass(): dust mass per entity (kg) */
} sk_dustmix_t;

/* ---- Sources: SourceSystem / NormalizedSource / GeometricSource / PointSource ------------------- */
enum sk_source_kind { SK_SRC_POINT = 1, SK_SRC_GEOMETRIC = 2 };
enum sk_geometry_kind {
    SK_GEOM_NONE = 0,
    SK_GEOM_SHELL = 1,          /* ShellGeometry.cpp:14-28,43-57 p = {rmin, rmax, exponent, _smin, _sdiff, _tmin, _tmax} */
    SK_GEOM_EXPDISK = 2,        /* ExpDiskGeometry.cpp:46-68   p = {hR, hz, Rmin, Rmax, zmax} */
    SK_GEOM_RING = 3,           /* RingGeometry.cpp:56-68      p = {R0, w, hz} + radial cdf table */
    SK_GEOM_SPIRAL_EXPDISK = 4  /* SpiralStructureGeometryDecorator.cpp:33-45,72-76 on ExpDisk:
                                   p = {hR, hz, Rmin, Rmax, zmax, m(arms), tan(pitch) = _tanp, R0, phi0, w, N(index),
                                        _cn} with the setup values of SpiralStructureGeometryDecorator.cpp:12-20 */
};
enum sk_sed_kind {
    SK_SED_TABULATED = 1, /* specificLuminosity by log-log interpolation of (lambda,p) (TabulatedSED) */
    SK_SED_BLACKBODY = 2  /* specificLuminosity = Planck(lambda,T)/Ltot (BlackBodySED.cpp:38-41) */
};
enum sk_bias_kind {
    SK_BIAS_NONE = 0,
    SK_BIAS_LOGUNIFORM = 1, /* DefaultWavelengthDistribution.cpp:12-40  range = {min,max} */
    SK_BIAS_OLIGO = 2       /* OligoWavelengthDistribution.cpp:13-38    discrete wavelengths */
};
#define SK_GEOM_MAX_PARAMS 12
typedef struct sk_source {
    int32_t kind;                /* sk_source_kind */
    int32_t geometry;            /* sk_geometry_kind (SK_SRC_GEOMETRIC only) */
    double luminosity;           /* Source::luminosity() (W) */
    double source_weight;        /* Source::sourceWeight() */
    double position[3];          /* PointSource position */
    double geom_params[SK_GEOM_MAX_PARAMS];
    int32_t geom_table_n;        /* RingGeometry::_Rv/_Xv size (0 if none) */
    int32_t sed_kind;            /* sk_sed_kind */
    const double* geom_table_x;  /* [geom_table_n] */
    const double* geom_table_P;  /* [geom_table_n] */
    int32_t sed_n;               /* size of the SED cdf tables */
    int32_t bias_kind;           /* sk_bias_kind */
    const double* sed_lambda;    /* [sed_n] SED::_lambdav */
    const double* sed_p;         /* [sed_n] normalised pdf _pv */
    const double* sed_P;         /* [sed_n] normalised cdf _Pv */
    double sed_temperature;      /* blackbody T (K) */
    double sed_norm;             /* blackbody: _Ltot of BlackBodySED.cpp:16 */
    double wavelength_bias;      /* NormalizedSource::_xi (forced to 1 for oligochromatic) */
    double bias_min, bias_max;   /* log-uniform bias distribution range */
    int32_t oligo_n;             /* number of discrete wavelengths */
    int32_t reserved;
    const double* oligo_lambda;  /* [oligo_n] */
    double oligo_probability;    /* OligoWavelengthDistribution::_probability */
    int32_t velocity_kind;       /* sk_velocity_kind: the bulk velocity of the source, Source::hasVelocity() -- the packet is
                                    launched at lambda (1 - k.v/c) (PhotonPacket::launch, PhotonPacket.cpp:33) */
    int32_t reserved2;
    double velocity[3];          /* SK_VEL_CONSTANT: the vector (m/s); SK_VEL_RADIAL / SK_VEL_CYLINDRICAL: {magnitude,
                                    unityRadius, exponent} of GeometricSource::velocityMagnitude() times the field */
} sk_source_t;
/* The velocity of a source at the launch position: PointSource velocityX/Y/Z (SpecialtySource.cpp) or
 * GeometricSource::velocityMagnitude() * velocityDistribution()->vector(r) (GeometricSource.cpp:66-82) for the vector fields
 * UnidirectionalVectorField (as a constant vector), RadialVectorField.cpp:19-37 and CylindricalVectorField.cpp:19-38. */
enum sk_velocity_kind { SK_VEL_NONE = 0, SK_VEL_CONSTANT = 1, SK_VEL_RADIAL = 2, SK_VEL_CYLINDRICAL = 3 };

/* ---- Instruments: DistantInstrument / SEDInstrument / FrameInstrument / FullInstrument --------- */
enum sk_instrument_kind { SK_INSTR_SED = 1, SK_INSTR_FRAME = 2, SK_INSTR_FULL = 3 };
typedef struct sk_instrument {
    int32_t kind;                   /* sk_instrument_kind */
    int32_t wavelength_grid;        /* index into the grids given to sk_engine_set_wavelength_grids */
    double inclination, azimuth, roll; /* rad (DistantInstrument.cpp:39-50) */
    double distance;                /* m */
    double radius;                  /* aperture radius (SEDInstrument / ApertureInstrument.cpp:24-43); 0 = none */
    int32_t num_pixels_x, num_pixels_y;
    double field_of_view_x, field_of_view_y, center_x, center_y; /* FrameInstrument.cpp:12-32 */
    int32_t record_components;      /* Instrument::recordComponents */
    int32_t num_scattering_levels;  /* Instrument::numScatteringLevels */
    int32_t record_statistics;      /* Instrument::recordStatistics: Sum w^k per SED bin and per frame pixel, with the
                                       contributions of one history to the same bin combined first
                                       (FluxRecorder.cpp:457-466, 962-1014) */
    int32_t reserved;
    double redshift;                /* observer-frame redshift z of the instrument's recorder: a packet is binned at
                                       lambda (1 + z) (FluxRecorder::setObserverFrameRedshift, FluxRecorder.cpp:123-130,
                                       309-310); 0 for the local universe */
} sk_instrument_t;

/* Detector array ids: the enum of SKIRT/core/FluxRecorder.cpp:26-56 without the polarisation entries. */
enum sk_component {
    SK_COMP_TOTAL = 0,
    SK_COMP_TRANSPARENT = 1,
    SK_COMP_PRIMARY_DIRECT = 2,
    SK_COMP_PRIMARY_SCATTERED = 3,
    SK_COMP_SECONDARY_DIRECT = 4,
    SK_COMP_SECONDARY_SCATTERED = 5,
    SK_COMP_SECONDARY_TRANSPARENT = 6,
    SK_COMP_PRIMARY_SCATTERED_LEVEL = 7 /* + level-1 */
};

/* ---- Secondary (dust) emission: DustSecondarySource + EquilibriumDustEmissionCalculator --------
 * The tables are what EquilibriumDustEmissionCalculator::precalculate (EquilibriumDustEmissionCalculator.cpp:18-93)
 * leaves behind for the single dust mix; they are setup data computed on the host side like the DustMix tables. */
typedef struct sk_secondary {
    int32_t emission_grid;       /* wavelength grid index of Configuration::dustEmissionWLG() */
    int32_t num_temperatures;    /* size of the temperature grid _Tv (1001, .cpp:55) */
    double spatial_bias;         /* SecondaryEmissionOptions::spatialBias (DustSecondarySource.cpp:118-146) */
    double wavelength_bias;      /* DustEmissionOptions::wavelengthBias (DustSecondarySource.cpp:526) */
    double bias_min, bias_max;   /* range of the log-uniform DefaultWavelengthDistribution = wavelength range of the
                                    dust emission grid (DustEmissionOptions.hpp:87-90) */
    const double* temperature;   /* [num_temperatures] _Tv */
    const double* planck_abs;    /* [num_temperatures] _planckabsvv[0] (.cpp:70-91) */
    const double* rf_sigma_abs;  /* [N_rf]   _rfsigmaabsvv[0]: sigma_abs resampled on the radiation field grid (.cpp:59) */
    const double* em_sigma_abs;  /* [N_em+2] _emsigmaabsvv[0]: sigma_abs on DisjointWavelengthGrid::extlambdav() of the
                                    emission grid (.cpp:62-65, DisjointWavelengthGrid.cpp:346-356) */
    const double* rf_cmb;        /* [N_rf] _Bcmbv: the CMB source term B_lambda(T_cmb (1 + z)) at the wavelengths of the radiation
                                    field grid, added to the mean intensity in the energy balance when DustEmissionOptions::
                                    includeHeatingByCMB (.cpp:37-44, 120-131); NULL = none */
} sk_secondary_t;

/* ---- Setup on the device: octree construction and medium-state sampling ----
 * A medium component as the tree policy and the density sampler see it: a normalised Geometry (density integrates to one)
 * times a total number of entities / a total mass (GeometricMedium.cpp:14-40).  p[] holds what Geometry::density needs:
 *   SK_GEOM_SHELL           {rmin, rmax, exponent, _A}                                   ShellGeometry.cpp:30-36
 *   SK_GEOM_EXPDISK         {hR, hz, Rmin, Rmax, zmax, _rho0}                            ExpDiskGeometry.cpp:32-42
 *   SK_GEOM_RING            {R0, w, hz, _A}                                              RingGeometry.cpp:39-43
 *   SK_GEOM_SPIRAL_EXPDISK  {hR, hz, Rmin, Rmax, zmax, _rho0, m, _tanp, R0, phi0, w, N, _cn}
 *                                                                     SpiralStructureGeometryDecorator.cpp:24-29,71-75 */
#define SK_DENSITY_MAX_PARAMS 16
typedef struct sk_density_geometry {
    int32_t geometry;  /* sk_geometry_kind */
    int32_t reserved;
    double number;     /* Medium::number(): number density = number * Geometry::density(r) */
    double mass;       /* Medium::mass():   mass density   = mass   * Geometry::density(r) */
    double p[SK_DENSITY_MAX_PARAMS];
} sk_density_geometry_t;

/* DensityTreePolicy (DensityTreePolicy.cpp:116-227), dust criteria; a criterion with value 0 is disabled. */
typedef struct sk_tree_policy {
    int32_t min_level, max_level;       /* TreePolicy::minLevel / maxLevel */
    int32_t num_samples;                /* SamplingOptions::numDensitySamples (DensityTreePolicy.cpp:76) */
    int32_t reserved;
    double max_dust_fraction;           /* delta = M_node / M_dust > maxDustFraction             (.cpp:192-196) */
    double max_dust_optical_depth;      /* tau = kappa * rho * diagonal > maxDustOpticalDepth    (.cpp:199-203) */
    double max_dust_density_dispersion; /* q = (rhomax-rhomin)/rhomax > maxDustDensityDispersion (.cpp:206-210) */
    double dust_kappa;                  /* DensityTreePolicy::_dustKappa (.cpp:88-97) */
} sk_tree_policy_t;

/* ---- Device-side event counters (the engine must count S, S_fwd, P_peel itself) -- */
typedef struct sk_counters {
    uint64_t packets;        /* histories launched with L > 0 */
    uint64_t forward_paths;  /* setExtinctionOpticalDepths calls */
    uint64_t forward_segments;
    uint64_t replay_segments;/* segments re-walked to the interaction point (engine-internal; not algorithmic) */
    uint64_t peel_paths;     /* getExtinctionOpticalDepth calls */
    uint64_t peel_segments;
    uint64_t scatterings;    /* simulateScattering calls */
    uint64_t rf_deposits;    /* MediumSystem::storeRadiationField calls */
    uint64_t detections;     /* FluxRecorder::detect calls that recorded */
    uint64_t fallbacks;      /* tree: top-down relocations after a failed neighbour link */
    uint64_t kernel_launches;/* kernels launched by sk_engine_launch_segment / prepare_secondary (host-side count) */
    uint64_t rounds;         /* rounds of the stage sequence over the bank */
    uint64_t pixel_overflows;/* per-pixel statistics: contributions recorded on their own because the pool of list chunks
                                ran dry (0 unless SK_PIX_POOL was set too small; then Sum w^k, k >= 2, is biased) */
    uint64_t reserved[3];
} sk_counters_t;

/* ---- life cycle of the engine object ---------------------------------------------------------- */
int sk_abi_version(void);
const char* sk_last_error(void);

/* Replaces nothing in the reference; allocates the per-GPU context the shim keeps next to
 * MonteCarloSimulation (one engine per CUDA device / rank). */
int sk_engine_create(const sk_config_t* config, sk_engine_t** out);
void sk_engine_destroy(sk_engine_t* e);

/* CartesianSpatialGrid::_xv/_yv/_zv (CartesianSpatialGrid.cpp:22-60); cell index m = k + Nz*j + Nz*Ny*i
 * (CartesianSpatialGrid.cpp:210-213).  xv has nx+1 ascending borders, etc. */
int sk_engine_set_grid_cartesian(sk_engine_t* e, int32_t nx, int32_t ny, int32_t nz, const double* xv,
                                 const double* yv, const double* zv);

/* TreeSpatialGrid::_nodev for an OctTreeNode tree (TreeSpatialGrid.cpp:33-49, OctTreeNode.cpp:22-35):
 * extent = {xmin,ymin,zmin,xmax,ymax,zmax} of the root; first_child[l] = node index of the first of
 * the 8 consecutive children of node l (child order OctTreeNode.cpp:22-35), or -1 when node l is a leaf.
 * Cell index m = rank of the leaf in node order (TreeSpatialGrid::_cellindexv).  The engine derives its
 * own device layout (lattice border tables, per-cell neighbour links) from this. */
int sk_engine_set_grid_octree(sk_engine_t* e, const double extent[6], int32_t num_nodes, const int32_t* first_child);

/* VoronoiMeshSnapshot (VoronoiMeshSnapshot.cpp:85-194, 491-730) as built by the reference's setup with the vendored voro++:
 * sites[3*m..] = Cell::position() of cell m; the neighbours of cell m are nbr_index[nbr_offset[m] .. nbr_offset[m+1]) in the
 * order of Cell::neighbors(): a cell index, or -1..-6 for the domain walls xmin,xmax,ymin,ymax,zmin,zmax
 * (VoronoiMeshSnapshot.cpp:1134-1143).  extent = the domain box.  The engine derives its own nearest-site search structure
 * (replacing the block lists and k-d trees of VoronoiMeshSnapshot.cpp:765-825, 1006-1040). */
int sk_engine_set_grid_voronoi(sk_engine_t* e, const double extent[6], int32_t num_cells, const double* sites,
                               const int64_t* nbr_offset, const int32_t* nbr_index);

/* VoronoiMeshSnapshot::buildMesh (VoronoiMeshSnapshot.cpp:491-730) on the device: the tessellation of the domain box by the
 * given sites -- what the reference obtains from the vendored voro++ (container::compute_cell per site; neighbors, volume,
 * vertices of every cell).  sites[3*m..] as in sk_engine_set_grid_voronoi, all strictly inside the domain and in the order the
 * cells are to be numbered (the reference sorts them by x and drops sites outside the domain or closer than 1e-12 of its
 * diagonal to another, VoronoiMeshSnapshot.cpp:500-540; the caller does that).  One thread builds one cell: the box around the
 * site is cut with the bisecting planes towards the other sites, visiting the blocks of a uniform search grid outwards, until
 * every unvisited site is farther away than twice the cell's largest vertex distance.  Leaves the engine in the state
 * sk_engine_set_grid_voronoi + sk_engine_set_voronoi_extents would, and keeps the cell volumes (MediumState::volume).  Sites in
 * degenerate position (more than three planes through a vertex, as in a regular lattice) are reported as SK_ERR_UNSUPPORTED --
 * the caller then falls back to its own tessellation.  *num_entries = total length of the neighbour lists. */
int sk_engine_build_voronoi(sk_engine_t* e, const double extent[6], int32_t num_sites, const double* sites,
                            uint64_t* num_entries);
/* The tessellation the engine holds after sk_engine_build_voronoi (any pointer may be NULL): nbr_offset[num_cells+1],
 * nbr_index[num_entries] (per cell the neighbours that share a face with it, then the domain walls -1..-6 that bound it),
 * volume[num_cells], boxes[6*num_cells]. */
int sk_engine_read_voronoi(sk_engine_t* e, int64_t* nbr_offset, int32_t* nbr_index, double* volume, double* boxes);

/* Optional, after sk_engine_set_grid_voronoi: the enclosing box of every Voronoi cell, boxes[6*m..] = {xmin,ymin,zmin,xmax,
 * ymax,zmax} of VoronoiMeshSnapshot::Cell (a Box: the bounding box of the cell's vertices, VoronoiMeshSnapshot.cpp:104-135).
 * Needed only for dust emission from a Voronoi grid: VoronoiMeshSnapshot::generatePosition(m) (.cpp:976-989) draws
 * Random::position(box) until the point is closest to site m among m's neighbours (isPointClosestTo, .cpp:848-856). */
int sk_engine_set_voronoi_extents(sk_engine_t* e, int32_t num_cells, const double* boxes);

/* MediumState number densities and volumes for a single medium component
 * (MediumState::numberDensity(m,0), MediumState::volume(m); MediumState.cpp:196-247). */
int sk_engine_set_medium(sk_engine_t* e, int32_t num_cells, const double* number_density, const double* volume);

/* DensityTreePolicy::constructTree (DensityTreePolicy.cpp:242-309) with OctTreeNode::createChildren (OctTreeNode.cpp:22-35),
 * on the device: level by level, every node of the level is evaluated by needsSubdivide (.cpp:116-227: num_samples
 * random positions in the node, Random::position(Box) Random.cpp:168-176, dust mass density summed over the media) and the
 * flagged nodes get 8 children appended in node order, so that the node list has the reference's breadth-first order.
 * The random positions come from the engine's counter-based generator (key = (seed, 0x54524545 "TREE"), counter =
 * (node index, draw)); the reference draws from its thread-local Mersenne twisters, so the two trees agree statistically,
 * not node by node.  Leaves the engine in the state sk_engine_set_grid_octree would (tables, neighbour links). */
int sk_engine_build_octree(sk_engine_t* e, const double extent[6], const sk_tree_policy_t* policy, int32_t num_media,
                           const sk_density_geometry_t* media, uint64_t* num_nodes, uint64_t* num_cells);
/* The first_child array of the octree the engine holds (same meaning as in sk_engine_set_grid_octree), e.g. for
 * TreeSpatialGridTopologyProbe (TreeSpatialGrid.cpp:232-251) or to rebuild the reference's TreeNode objects. */
int sk_engine_read_octree(sk_engine_t* e, int32_t* first_child);

/* MediumSystem::setupSelfAfter's cell loop (MediumSystem.cpp:286-330) for one geometric medium on a Cartesian or octree
 * grid: volume(m) = SpatialGrid::volume(m), numberDensity(m) = PropertySampler::density (MediumSystem.cpp:80-106): the
 * density at the cell centre for num_samples = 1, else the mean over num_samples random positions in the cell
 * (TreeSpatialGrid::randomPositionInCell, TreeSpatialGrid.cpp:125-128).  Generator key = (seed, 0x43454c4c "CELL"),
 * counter = (cell index, draw).  Replaces sk_engine_set_medium. */
int sk_engine_sample_medium(sk_engine_t* e, const sk_density_geometry_t* medium, int32_t num_samples);
/* The same cell loop for a ParticleMedium: the smoothed-particle density of ParticleSnapshot::density(Position)
 * (ParticleSnapshot.cpp:248-258) -- sum over the particles, in ascending index, of W(|r - c_m| / h_m) M_m / h_m^3 with the
 * CubicSplineSmoothingKernel (CubicSplineSmoothingKernel.cpp:39-47) -- times density_scale (ImportedMedium::numberDensity,
 * ImportedMedium.cpp:198-203: 1/mu for a snapshot that holds masses, times the mass fraction), averaged over num_samples random
 * positions in the cell: Random::position(box) for Cartesian and octree cells, VoronoiMeshSnapshot::generatePosition(m)
 * (VoronoiMeshSnapshot.cpp:976-989) for Voronoi cells, which needs the cells' extents; num_samples = 1 takes the centre of the
 * box (not available for Voronoi cells, whose central position is the centroid).  particles[5*m..] = x y z h M.  Volumes: box
 * volumes, or the volumes of a tessellation built by sk_engine_build_voronoi.  Same generator key and counter as
 * sk_engine_sample_medium. */
int sk_engine_sample_medium_particles(sk_engine_t* e, int32_t num_particles, const double* particles, double density_scale,
                                      int32_t num_samples);
/* MediumState::numberDensity(m,0) and MediumState::volume(m) as the engine holds them (either array may be NULL). */
int sk_engine_read_medium(sk_engine_t* e, double* number_density, double* volume);

int sk_engine_set_dustmix(sk_engine_t* e, const sk_dustmix_t* mix);

/* Kinematics (f4 of SURVEY.md section 8): MediumState::bulkVelocity(m) (MediumSystem.cpp:330-365), velocity[3*m + c] in m/s,
 * after the medium state; NULL (or num_cells = 0) returns to media at rest.  With moving media -- Configuration::
 * hasMovingMedia(), and here also with moving sources alone -- the wavelength a cell perceives differs from cell to cell,
 * lambda / (1 - k.v_m/c) (PhotonPacket::perceivedWavelength, PhotonPacket.cpp:133-151), and every use of a cross section
 * follows it: the optical depths of the forward and peel-off paths look the sections up per segment (the "spatially variable
 * cross sections" branches, MediumSystem.cpp:888-900, 958-972, 1242-1258), the radiation field is binned at the perceived
 * wavelength with the perceived luminosity (MonteCarloSimulation.cpp:667-691), albedo, peel-off weights and the scattering
 * component use the wavelength perceived in the interaction cell (MediumSystem.cpp:667-693), a scattered or peeled-off packet
 * leaves at lambda_perceived (1 - k_new.v_m/c) (PhotonPacket::scatter / launchScatteringPeelOff, PhotonPacket.cpp:89-122), and
 * a packet emitted by a dust cell carries the cell's bulk velocity (DustSecondarySource.cpp:562-580).  The caller passes the
 * path length bias the configuration ends up with (0 with moving media and forced scattering, Configuration.cpp:492-498). */
int sk_engine_set_velocities(sk_engine_t* e, int32_t num_cells, const double* velocity);

/* Several medium components, each with its own material mix (Configuration::hasMultipleConstantSectionMedia; f4 of SURVEY.md
 * section 8): the medium state MediumState::numberDensity(m,h) for h = 0..num_media-1, number_density[h*num_cells + m], and
 * the mix of every component.  All dust mixes of one simulation share the wavelength grid of their property tables
 * (DustMix.cpp:52-98 derives it from the configuration alone), so lambda_border must be the same in all of them.  What changes
 * on the path: optical depths sum sigma_h n_h over the components (MediumSystem.cpp:874-885, 1222-1240, 1012-1040), the
 * albedo is sum k_sca / sum k_ext in the interaction cell (:678-693), a scattering peel-off is the sum of the components'
 * phase functions weighted with their scattering opacities (:697-767), the scattering component is drawn from those weights
 * with one extra deviate (:796-823), and the absorbed luminosity and the dust emission spectrum sum over the components, each
 * with the equilibrium temperature of its own mix (:1317-1356, 1452-1476).  sk_engine_set_medium / sk_engine_set_dustmix are
 * the num_media = 1 forms.  Up to SK_MAX_MEDIA components.  With explicit absorption the walks are in scattering optical depth and
 * the absorption optical depth is accumulated next to it and interpolated at the interaction point (MediumSystem.cpp:937-955,
 * 1112-1150). */
/* May be called again on a configured engine with the same grid -- a dynamic medium state whose recipes have changed densities
 * between two segments (MediumSystem::updateDynamicStateRecipes, MediumSystem.cpp:1498-1558; the primary / merged iteration
 * loops of MonteCarloSimulation.cpp:266-330, 407-496): grid links, dust tables, radiation field and detector arrays stay; the
 * new medium state is at rest until sk_engine_set_velocities is called again. */
int sk_engine_set_media(sk_engine_t* e, int32_t num_cells, int32_t num_media, const double* number_density,
                        const double* volume);
int sk_engine_set_dustmixes(sk_engine_t* e, int32_t num_media, const sk_dustmix_t* mixes);

/* All wavelength grids used by instruments, the radiation field and dust emission, addressed by index.
 * rf_grid = index of Configuration::radiationFieldWLG() or -1 when no radiation field is stored. */
int sk_engine_set_wavelength_grids(sk_engine_t* e, int32_t n, const sk_wavelength_grid_t* grids, int32_t rf_grid);

/* SourceSystem::setupSelfAfter (SourceSystem.cpp:14-41): sources + SourceSystem::sourceBias(). */
int sk_engine_set_sources(sk_engine_t* e, int32_t n, const sk_source_t* sources, double source_bias);

/* InstrumentSystem::instruments(); allocates and zeroes the detector arrays
 * (FluxRecorder::finalizeConfiguration, FluxRecorder.cpp:185-300).  has_medium_emission mirrors
 * Instrument.cpp:23 (secondary component arrays exist only then). */
int sk_engine_set_instruments(sk_engine_t* e, int32_t n, const sk_instrument_t* instruments, int32_t has_medium_emission);

int sk_engine_set_secondary(sk_engine_t* e, const sk_secondary_t* sec);
/* The same for several dust components: sec[h] holds the calculator tables of component h's mix (planck_abs, rf_sigma_abs,
 * em_sigma_abs); the grid, the temperature grid and the bias settings must agree in all entries. */
int sk_engine_set_secondary_media(sk_engine_t* e, int32_t num_media, const sk_secondary_t* sec);

/* Zeroes all detector and statistics arrays (what FluxRecorder::finalizeConfiguration leaves behind,
 * FluxRecorder.cpp:185-300) so that one engine can run several simulations back to back. */
int sk_engine_clear_instruments(sk_engine_t* e);

/* MediumSystem::clearRadiationField(primary) (MediumSystem.cpp:1279-1290). */
int sk_engine_clear_rf(sk_engine_t* e, int32_t primary);

/* SourceSystem::prepareForLaunch(numPackets) (SourceSystem.cpp:75-97). */
int sk_engine_prepare_primary(sk_engine_t* e, uint64_t num_packets);

/* SecondarySourceSystem::prepareForLaunch + DustSecondarySource::prepareLuminosities/preparePacketMap
 * (SecondarySourceSystem.cpp:84-126, DustSecondarySource.cpp:26-146) and, for every emitting cell at once, the
 * emission spectrum and its cumulative distribution that the reference computes lazily per thread
 * (DustCellEmission::calculateIfNeeded, DustSecondarySource.cpp:187-285; MediumSystem::dustEmissionSpectrum/
 * meanIntensity, MediumSystem.cpp:1370-1380,1466-1476; EquilibriumDustEmissionCalculator::emissivity, .cpp:120-150);
 * computed on the device from the current radiation field rf1+rf2.  Returns the total dust luminosity (W) in
 * *luminosity; a zero luminosity is reported as SK_OK with *luminosity = 0 (the caller skips the segment like
 * MonteCarloSimulation.cpp:156-159). */
int sk_engine_prepare_secondary(sk_engine_t* e, uint64_t num_packets, double* luminosity);

/* Sharding of a segment over several engines (GPUs, ranks) without a chunk server (the reference hands out chunks of
 * history indices dynamically, MultiHybridParallel.cpp:26-144, ChunkMaker.cpp:14-50): after this call the engine runs,
 * of the range [first, first+count) given to sk_engine_run_segment, only the histories i with
 * ((i - first) / block) % num_parts == part, i.e. every num_parts-th block of `block` consecutive histories (block a
 * power of two).  All engines of a run are then given the SAME range.  Interleaved blocks balance the load where
 * contiguous shares do not: the histories of a secondary emission segment are ordered by cell
 * (DustSecondarySource.cpp:133-145), and a contiguous share is a region of the model with its own path lengths.  The
 * random streams are keyed by history index, so the tallies do not depend on the sharding.  num_parts = 1 (the default)
 * runs the whole range. */
int sk_engine_set_history_interleave(sk_engine_t* e, uint64_t block, uint32_t num_parts, uint32_t part);

/* THE hot path: performLifeCycle(firstIndex, numIndices, primary, peel, store)
 * (MonteCarloSimulation.cpp:538-613) for histories [first, first+count) of the segment prepared by
 * sk_engine_prepare_*; stream_id distinguishes the random streams of successive segments (Philox key
 * word 1).  Returns after the device has finished (same semantics as Parallel::call returning) and
 * includes InstrumentSystem::flush() (FluxRecorder.cpp:472-480). */
int sk_engine_run_segment(sk_engine_t* e, uint64_t first, uint64_t count, int32_t primary, int32_t peel,
                          int32_t store, uint32_t stream_id);
/* The same without the final stream synchronisation and without reading the per-stage timings back: the host drives the
 * rounds of the stage sequence (it looks at the census of the bank one round behind the device, so the device never waits
 * for it) and returns when the last round has been ENQUEUED and the census says the bank is empty; work that the caller
 * enqueues on sk_engine_cuda_stream() afterwards -- an NCCL all-reduce of the tallies -- is ordered behind the segment
 * without a host synchronisation.  Pair with sk_engine_synchronize(). */
int sk_engine_launch_segment(sk_engine_t* e, uint64_t first, uint64_t count, int32_t primary, int32_t peel,
                             int32_t store, uint32_t stream_id);
int sk_engine_synchronize(sk_engine_t* e);
/* Device time (ms) of the last segment (all its stage kernels), measured with CUDA events on the engine's stream. */
int sk_engine_last_kernel_ms(sk_engine_t* e, float* ms);
/* Device time (ms) of the last segment per stage kernel, summed over the rounds (CUDA events around every launch):
 * out[SK_STAGE_*]. */
enum sk_stage { SK_STAGE_ADVANCE = 0, SK_STAGE_LAUNCH, SK_STAGE_PEEL_SETUP, SK_STAGE_DETECT, SK_STAGE_SAMPLE, SK_STAGE_TRACE_FORWARD,
                SK_STAGE_TRACE_INTERACTION, SK_STAGE_TRACE_PEEL, SK_STAGE_COUNT };
int sk_engine_last_stage_ms(sk_engine_t* e, float out[SK_STAGE_COUNT]);

/* MediumSystem::communicateRadiationField(primary) for the single-process case: _rf2 = _rf2c
 * (MediumSystem.cpp:1304-1313).  Across GPUs the caller all-reduces the device buffer first
 * (sk_engine_device_buffer + NCCL), exactly where the reference calls ProcessManager::sumToAll. */
int sk_engine_communicate_rf(sk_engine_t* e, int32_t primary);

/* MediumSystem::totalDustAbsorbedLuminosity(primary) (MediumSystem.cpp:1317-1356). */
int sk_engine_absorbed_luminosity(sk_engine_t* e, int32_t primary, double* out);

/* Outputs. which: 0 = _rf1, 1 = _rf2, 2 = _rf2c; out[m*Nrf + ell] (Table<2> row-major, MediumSystem.hpp:883-885). */
int sk_engine_read_rf(sk_engine_t* e, int32_t which, double* out);
/* FluxRecorder::_sed[component][ell] / _ifu[component][l + ell*Npix] (FluxRecorder.cpp:433).  For
 * SK_COMP_TOTAL with recordComponents the engine returns the sum the reference forms in
 * FluxRecorder::calibrateAndWrite (FluxRecorder.cpp:540-570). */
int sk_engine_read_sed(sk_engine_t* e, int32_t instrument, int32_t component, double* out);
int sk_engine_read_ifu(sk_engine_t* e, int32_t instrument, int32_t component, double* out);
/* FluxRecorder::_wsed[k][ell], k = 0..4 (FluxRecorder.cpp:58-62). */
int sk_engine_read_sed_stats(sk_engine_t* e, int32_t instrument, int32_t k, double* out);
/* FluxRecorder::_wifu[k][l + ell*Npix], k = 0..4: the same statistics per frame pixel (FluxRecorder.cpp:990-1013). */
int sk_engine_read_ifu_stats(sk_engine_t* e, int32_t instrument, int32_t k, double* out);
int sk_engine_counters(sk_engine_t* e, sk_counters_t* out, int32_t reset);

/* Raw device buffers so that the rank's communicator (NCCL via torch.distributed in bench.py, or the
 * shim's own ncclAllReduce) can reduce tallies in place: which = 0 rf1, 1 rf2, 2 rf2c, 3 = all
 * detector arrays of all instruments (one contiguous block), 4 = all statistics arrays.  The radiation field buffers are
 * wavelength-major on the device, [ell*Ncells + m] (the same on every rank, so element-wise reductions are unaffected);
 * sk_engine_read_rf hands them out in the reference's [m*Nrf + ell] layout. */
int sk_engine_device_buffer(sk_engine_t* e, int32_t which, void** device_ptr, uint64_t* num_doubles);
/* Diagnostic for bench.py's second roofline: the measured rate (records/s) at which this device serves chains of
 * dependent, randomly scattered 32-byte record fetches out of a table of num_records records at full occupancy -- the
 * memory access pattern of the crossing loop (one cell record per crossing, the next cell index comes out of it)
 * without its arithmetic.  No counterpart in the reference. */
int sk_engine_measure_gather_peak(sk_engine_t* e, int32_t num_records, double* records_per_s);
/* The cudaStream_t all engine work is enqueued on (so that a caller can order its collectives and its CUDA
 * events after the life-cycle kernel without a host synchronisation). */
int sk_engine_cuda_stream(sk_engine_t* e, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* SK_ENGINE_H */
