"""Small seeded models shared by the oracle tests (CPU) and the parity tests (GPU)."""
import math

import numpy as np

from skirt9_b200 import abi, configs
from skirt9_b200 import host as H

DEG = math.pi / 180.0


def small_cartesian(num_packets=20000, seed=3, **kw):
    return configs.cfg1(num_packets=num_packets, seed=seed, num_density_samples=4, **kw)


def small_octree(num_packets=20000, seed=5, **kw):
    args = dict(max_level=6, max_dust_fraction=1e-4, num_pixels=32, num_wavelengths=8)
    args.update(kw)
    return configs.cfg2(num_packets=num_packets, seed=seed, **args)


def small_octree_engine_setup(num_packets=20000, seed=5, **kw):
    """small_octree with the tree and the densities built on the engine's side (sk_engine_build_octree /
    sk_engine_sample_medium) instead of by the numpy mirror."""
    sim = small_octree(num_packets=num_packets, seed=seed, **kw)
    sim.deviceSetup = True
    return sim


def two_sources_three_instruments(num_packets=20000, seed=11, force=True):
    """Point + shell sources, SED (with aperture) + two frames that share one observer, scattering levels,
    statistics, strongly forward-scattering dust (exercises the averaged HG peel-off), ragged 7x5x3 grid."""
    pc = H.PC
    mix = H.MeanListDustMix([0.1e-6, 1e-6, 10e-6], [2000.0, 1000.0, 100.0], [0.7, 0.6, 0.4], [0.97, 0.96, 0.2])
    medium = H.GeometricMedium(H.ShellGeometry(0.05 * pc, 0.9 * pc, 1.0), mix, opticalDepth=3.0, wavelength=0.55e-6)
    grid = H.CartesianSpatialGrid(-pc, pc, -0.8 * pc, 0.9 * pc, -0.5 * pc, 0.7 * pc, 7, 5, 3)
    s1 = H.PointSource((0.1 * pc, -0.2 * pc, 0.05 * pc), H.BlackBodySED(8000.0), luminosity=2.0 * H.LSUN)
    s2 = H.GeometricSource(H.ShellGeometry(0.2 * pc, 1.5 * pc, 2.0), H.BlackBodySED(3000.0), luminosity=1.0 * H.LSUN,
                           sourceWeight=2.0)
    wlg = H.LogWavelengthGrid(0.2e-6, 5e-6, 6)
    i1 = H.SEDInstrument(instrumentName="sed", distance=1e6 * pc, inclination=30 * DEG, azimuth=40 * DEG,
                         radius=0.6 * pc, recordComponents=True, numScatteringLevels=2, recordStatistics=True)
    i2 = H.FrameInstrument(instrumentName="f1", distance=1e6 * pc, inclination=100 * DEG, azimuth=-20 * DEG,
                           roll=15 * DEG, fieldOfViewX=2 * pc, numPixelsX=9, fieldOfViewY=1.5 * pc, numPixelsY=7,
                           centerX=0.1 * pc, recordComponents=False)
    i3 = H.FullInstrument(instrumentName="f2", distance=1e6 * pc, inclination=100 * DEG, azimuth=-20 * DEG,
                          roll=15 * DEG, fieldOfViewX=1 * pc, numPixelsX=4, fieldOfViewY=1 * pc, numPixelsY=4,
                          recordComponents=True, recordStatistics=True)
    return H.MonteCarloSimulation(sources=[s1, s2], medium=medium, grid=grid, instruments=[i1, i2, i3],
                                  numPackets=num_packets, minWavelength=0.15e-6, maxWavelength=8e-6,
                                  defaultWavelengthGrid=wlg, storeRadiationField=force,
                                  radiationFieldWLG=H.LogWavelengthGrid(0.15e-6, 8e-6, 5) if force else None,
                                  forceScattering=force, numDensitySamples=3, seed=seed)


def compare_engines(sim, a, b, rtol=1e-9):
    """Asserts that two engines that ran the same simulation agree: event counters exactly, tallies to rtol."""
    ca, cb = a.counters(), b.counters()
    for key in ("packets", "forward_paths", "forward_segments", "peel_paths", "peel_segments", "scatterings",
                "rf_deposits", "detections"):
        assert ca[key] == cb[key], (key, ca[key], cb[key])
    for j, ins in enumerate(sim.instruments):
        comps = [abi.SK_COMP_TOTAL]
        if ins.recordComponents:
            comps += [abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED]
            comps += [abi.SK_COMP_PRIMARY_SCATTERED_LEVEL + k for k in range(ins.numScatteringLevels)]
            if sim.dustEmissionWLG is not None:
                comps += [abi.SK_COMP_SECONDARY_DIRECT, abi.SK_COMP_SECONDARY_SCATTERED,
                          abi.SK_COMP_SECONDARY_TRANSPARENT]
        for c in comps:
            if ins.kind in (abi.SK_INSTR_SED, abi.SK_INSTR_FULL):
                x, y = a.read_sed(j, c), b.read_sed(j, c)
                np.testing.assert_allclose(x, y, rtol=rtol, atol=1e-300, err_msg=f"sed instr {j} comp {c}")
            if ins.kind in (abi.SK_INSTR_FRAME, abi.SK_INSTR_FULL):
                x, y = a.read_ifu(j, c), b.read_ifu(j, c)
                np.testing.assert_allclose(x, y, rtol=rtol, atol=rtol * max(y.max(), 1e-300),
                                           err_msg=f"ifu instr {j} comp {c}")
        if ins.recordStatistics and ins.kind in (abi.SK_INSTR_SED, abi.SK_INSTR_FULL):
            x, y = a.read_sed_stats(j), b.read_sed_stats(j)
            np.testing.assert_allclose(x, y, rtol=1e-8, atol=1e-300, err_msg=f"stats instr {j}")
        if ins.recordStatistics and ins.kind in (abi.SK_INSTR_FRAME, abi.SK_INSTR_FULL):
            x, y = a.read_ifu_stats(j), b.read_ifu_stats(j)
            np.testing.assert_array_equal(x[0], y[0], err_msg=f"pixel counts instr {j}")   # histories per pixel: integers
            for k in range(1, 5):
                np.testing.assert_allclose(x[k], y[k], rtol=1e-8, atol=1e-8 * y[k].max(), err_msg=f"pixel stats {k} instr {j}")
    if sim.storeRadiationField:
        x, y = a.read_rf(0), b.read_rf(0)
        np.testing.assert_allclose(x, y, rtol=rtol, atol=rtol * y.max())
        la, lb = a.absorbed_luminosity(True), b.absorbed_luminosity(True)
        assert abs(la - lb) <= 1e-9 * abs(lb)
        if sim.dustEmissionWLG is not None:
            for which in (1, 2):
                x, y = a.read_rf(which), b.read_rf(which)
                np.testing.assert_allclose(x, y, rtol=rtol, atol=rtol * max(y.max(), 1e-300), err_msg=f"rf {which}")


def small_dust_emission(num_packets=20000, seed=7, **kw):
    args = dict(max_level=5, max_dust_fraction=1e-3, num_sed_wavelengths=12, max_secondary_iterations=3)
    args.update(kw)
    return configs.cfg4(num_packets=num_packets, seed=seed, **args)


def small_voronoi(num_packets=20000, seed=11, num_sites=1500, **kw):
    """Random sites concentrated towards a disk, smooth disk density evaluated at the sites (cfg5-like, no fixture)."""
    pc = H.PC
    rng = np.random.default_rng(99)
    R = np.minimum(rng.gamma(2.0, 3000.0, size=num_sites), 15000.0)
    phi = rng.uniform(0, 2 * np.pi, size=num_sites)
    z = np.clip(rng.laplace(0.0, 300.0, size=num_sites), -1900.0, 1900.0)
    sites = np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1) * pc
    return configs.cfg5(sites, num_packets=num_packets, seed=seed, num_pixels=16, **kw)


def small_voronoi_dust_emission(num_packets=20000, seed=17, num_sites=1200, max_secondary_iterations=2):
    """cfg4's physics (hot point source in an r^-2 dust shell, dust emission with iterations) on a Voronoi grid of random
    sites concentrated towards the centre: SecondarySourceSystem launching from Voronoi cells
    (VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989)."""
    pc = H.PC
    rng = np.random.default_rng(5)
    r = 0.98 * rng.uniform(0.0, 1.0, size=num_sites) ** 1.5
    mu = rng.uniform(-1.0, 1.0, size=num_sites)
    phi = rng.uniform(0.0, 2 * np.pi, size=num_sites)
    st = np.sqrt(1.0 - mu * mu)
    sites = np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * mu], axis=1) * pc
    sim = configs.cfg4(num_packets=num_packets, seed=seed, num_sed_wavelengths=12,
                       max_secondary_iterations=max_secondary_iterations)
    sim.grid = H.VoronoiMeshSpatialGrid(-pc, pc, -pc, pc, -pc, pc, sites)
    sim.numDensitySamples = 1
    return sim


def tabulated_sed_ring_source_high_g(num_packets=20000, seed=13):
    """A ring-shaped source with a tabulated (ListSED) spectrum next to a point source, and a strongly forward-scattering
    dust mix (g = 0.97 > 0.95: the peel-off uses the +-4 degree averaged phase function, DustMix.cpp:395-445)."""
    pc = H.PC
    mix = H.MeanListDustMix([0.1e-6, 1e-6, 10e-6], [2000.0, 800.0, 60.0], [0.7, 0.6, 0.4], [0.97, 0.96, 0.3])
    medium = H.GeometricMedium(H.ShellGeometry(0.05 * pc, 1.0 * pc, 1.0), mix, opticalDepth=2.0, wavelength=0.55e-6)
    grid = H.PolicyTreeSpatialGrid(-pc, pc, -pc, pc, -pc, pc, H.DensityTreePolicy(2, 5, 2e-3))
    sed = H.ListSED([0.12e-6, 0.3e-6, 0.8e-6, 2e-6, 9e-6], [0.2, 1.0, 3.0, 1.5, 0.1])
    s1 = H.GeometricSource(H.RingGeometry(0.5 * pc, 0.1 * pc, 0.05 * pc), sed, luminosity=2.0 * H.LSUN)
    s2 = H.PointSource((0.1 * pc, 0.0, -0.2 * pc), H.BlackBodySED(8000.0), luminosity=1.0 * H.LSUN)
    wlg = H.LogWavelengthGrid(0.15e-6, 8e-6, 7)
    i1 = H.FullInstrument(instrumentName="f", distance=1e6 * pc, inclination=75 * DEG, azimuth=10 * DEG, fieldOfViewX=2.4 * pc,
                          numPixelsX=12, fieldOfViewY=2.4 * pc, numPixelsY=12, recordComponents=True, numScatteringLevels=3,
                          recordStatistics=True)
    return H.MonteCarloSimulation(sources=[s1, s2], medium=medium, grid=grid, instruments=[i1], numPackets=num_packets,
                                  minWavelength=0.15e-6, maxWavelength=8e-6, defaultWavelengthGrid=wlg,
                                  storeRadiationField=True, radiationFieldWLG=H.LogWavelengthGrid(0.15e-6, 8e-6, 5),
                                  numDensitySamples=3, seed=seed)


def long_histories_many_pixels(num_packets=1500, seed=21):
    """Histories that reach hundreds of distinct frame pixels: nearly conservative scattering (albedo 0.995) in an optically
    thick sphere, no path-length bias, a weight reduction of 1e4 before termination, a 64 x 64 frame with per-pixel statistics.  The per-history
    pixel list (FluxRecorder's ContributionList, FluxRecorder.cpp:990-1013) then is far longer than the SK_PIX_K entries
    that fit in a bank slot and continues in chunks from the pool."""
    pc = H.PC
    mix = H.MeanListDustMix([0.1e-6, 1e-6], [1000.0, 1000.0], [0.995, 0.995], [0.2, 0.2])
    medium = H.GeometricMedium(H.ShellGeometry(1e-4 * pc, 1.0 * pc, 0.0), mix, opticalDepth=100.0, wavelength=0.55e-6)
    grid = H.CartesianSpatialGrid(-pc, pc, -pc, pc, -pc, pc, 12, 12, 12)
    src = H.PointSource((0.0, 0.0, 0.0), H.BlackBodySED(5000.0), luminosity=1.0 * H.LSUN)
    instr = H.FullInstrument(instrumentName="i60", distance=1e6 * pc, inclination=60 * DEG, fieldOfViewX=2 * pc,
                             numPixelsX=64, fieldOfViewY=2 * pc, numPixelsY=64, recordComponents=True, recordStatistics=True)
    sim = H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr], numPackets=num_packets,
                                 oligoWavelengths=[0.55e-6], storeRadiationField=False, numDensitySamples=4, seed=seed)
    sim.pathLengthBias = 0.0   # (the bias weights p/q would end the histories after a dozen scatterings)
    return sim


def with_second_component(sim, kind="disk"):
    """Adds a second dust component with its own, greyer and more isotropically scattering mix to a model (the
    several-media paths of MediumSystem.cpp:678-823, 874-885, 1012-1040, 1222-1240, 1452-1476)."""
    pc = H.PC
    lam = sim.medium.mix.inlam
    n = len(lam)
    kappa = np.linspace(1800.0, 300.0, n)
    albedo = np.linspace(0.85, 0.35, n)
    g = np.linspace(0.35, 0.0, n)
    mix2 = H.MeanListDustMix(lam, kappa, albedo, g)
    scale = abs(sim.grid.extent[3]) if hasattr(sim.grid, "extent") else pc
    if kind == "disk":
        geom = H.ExpDiskGeometry(0.25 * scale, 0.05 * scale, 0.0, scale, 0.1 * scale)
    else:
        geom = H.ShellGeometry(0.2 * scale, 0.9 * scale, 0.0)
    sim.extraMedia = [H.GeometricMedium(geom, mix2, opticalDepth=0.5 * sim.medium.tau, wavelength=sim.medium.norm_wavelength)]
    return sim


def with_kinematics(sim, source=True, media=True, speed=6e6):
    """Gives a panchromatic model moving sources and / or moving media (PhotonPacket.cpp:133-151; Configuration::
    hasMovingSources / hasMovingMedia): the first medium component expands radially, a second one rotates about the z axis;
    a point source moves along a fixed direction, a geometric source expands.  `speed` of 6e6 m/s = 0.02 c shifts a wavelength
    over two dozen points of the dust property tables (1000 per dex)."""
    pc = H.PC
    scale = abs(sim.grid.extent[3]) if hasattr(sim.grid, "extent") else pc
    if media:
        for h, md in enumerate(sim.media):
            md.velocityMagnitude = speed * (1.0 if h == 0 else -0.7)
            md.velocityDistribution = H.RadialVectorField(scale, 1.0) if h == 0 else H.CylindricalVectorField(0.3 * scale, -0.5)
    if source:
        for s in sim.sources:
            if isinstance(s, H.PointSource):
                s.velocity = (0.8 * speed, -0.5 * speed, 0.3 * speed)
            else:
                s.velocityMagnitude = 0.6 * speed
                s.velocityDistribution = H.RadialVectorField(0.5 * scale, 0.5)
    return sim


def small_dynamic_state(num_packets=20000, seed=9):
    """cfg16d in small: a ClearDensityRecipe carves a cavity around a point source in primary emission iterations (packet
    ramp), merged primary + secondary iterations follow, then the regular segments (MonteCarloSimulation.cpp:266-330, 407-496)."""
    pc = H.PC
    mix = H.MeanListDustMix([0.05e-6, 0.55e-6, 10e-6, 100e-6, 2000e-6], [5000.0, 1000.0, 100.0, 5.0, 0.01],
                            [0.4, 0.6, 0.2, 0.01, 0.0001], [0.6, 0.5, 0.05, 0.0, 0.0])
    medium = H.GeometricMedium(H.ShellGeometry(0.02 * pc, 1.0 * pc, 1.0), mix, opticalDepth=6.0, wavelength=0.55e-6)
    grid = H.CartesianSpatialGrid(-pc, pc, -pc, pc, -pc, pc, 11, 11, 11)
    src = H.PointSource((0.0, 0.0, 0.0), H.BlackBodySED(10000.0), luminosity=1e4 * H.LSUN)
    instr = H.SEDInstrument(instrumentName="sed", distance=1e6 * pc, inclination=60 * DEG, recordComponents=True, recordStatistics=True)
    return H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr], numPackets=num_packets,
                                  minWavelength=0.1e-6, maxWavelength=20e-6,
                                  defaultWavelengthGrid=H.LogWavelengthGrid(0.1e-6, 1000e-6, 20), storeRadiationField=True,
                                  radiationFieldWLG=H.LogWavelengthGrid(0.1e-6, 1000e-6, 20),
                                  dustEmissionWLG=H.LogWavelengthGrid(1e-6, 1000e-6, 20), iterateSecondaryEmission=True,
                                  minSecondaryIterations=1, maxSecondaryIterations=4, secondaryIterationPacketsMultiplier=0.5,
                                  clearDensityThreshold=100.0, iteratePrimaryEmission=True, includePrimaryEmission=True,
                                  minPrimaryIterations=1, maxPrimaryIterations=6, primaryIterationPacketsMultiplier=0.5,
                                  primaryIterationInitialPacketsFraction=0.5, primaryIterationPacketsRamp=1.5,
                                  numDensitySamples=4, seed=seed)
