"""The workloads of BASELINE.json `configs`, written with the host mirror classes."""
from __future__ import annotations

import math

from . import host as H

DEG = math.pi / 180.0


def cfg1(num_packets=1e6, seed=0, record_statistics=False, num_density_samples=100):
    """point source in a uniform dust sphere, 32^3 Cartesian grid, one wavelength, RF stored."""
    pc = H.PC
    mix = H.MeanListDustMix([0.1e-6, 1e-6], [1000.0, 1000.0], [0.6, 0.6], [0.5, 0.5])
    medium = H.GeometricMedium(H.ShellGeometry(1e-4 * pc, 1.0 * pc, 0.0), mix, opticalDepth=2.0, wavelength=0.55e-6)
    grid = H.CartesianSpatialGrid(-pc, pc, -pc, pc, -pc, pc, 32, 32, 32)
    src = H.PointSource((0.0, 0.0, 0.0), H.BlackBodySED(5000.0), luminosity=1.0 * H.LSUN)
    instr = H.FullInstrument(instrumentName="i60", distance=1e6 * pc, inclination=60 * DEG, fieldOfViewX=2 * pc,
                             numPixelsX=64, fieldOfViewY=2 * pc, numPixelsY=64, recordComponents=True,
                             recordStatistics=record_statistics)
    return H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr],
                                  numPackets=num_packets, oligoWavelengths=[0.55e-6], storeRadiationField=True,
                                  numDensitySamples=num_density_samples, seed=seed)


def cfg2(num_packets=1e8, max_level=9, max_dust_fraction=3.5e-6, seed=0, num_pixels=256, num_wavelengths=50,
         record_statistics=False, min_level=3):
    """dusty spiral galaxy, octree of ~1e6 cells (933 059 in the reference at these defaults),
    50 wavelength bins, 256^2 FullInstrument, no radiation field."""
    pc = H.PC
    disk = H.ExpDiskGeometry(4000 * pc, 350 * pc, 0.0, 20000 * pc, 2000 * pc)
    spiral = H.SpiralStructureGeometryDecorator(disk, numArms=2, pitchAngle=20 * DEG, radiusZeroPoint=4000 * pc,
                                                phaseZeroPoint=0.0, perturbationWeight=0.5, index=1)
    src = H.GeometricSource(spiral, H.BlackBodySED(6000.0), luminosity=1e10 * H.LSUN)
    mix = H.MeanListDustMix([0.1e-6, 0.55e-6, 10e-6], [3000.0, 1000.0, 50.0], [0.5, 0.6, 0.3], [0.6, 0.5, 0.1])
    medium = H.GeometricMedium(H.RingGeometry(6000 * pc, 2000 * pc, 200 * pc), mix, opticalDepth=1.0,
                               wavelength=0.55e-6)
    grid = H.PolicyTreeSpatialGrid(-20000 * pc, 20000 * pc, -20000 * pc, 20000 * pc, -2000 * pc, 2000 * pc,
                                   H.DensityTreePolicy(min_level, max_level, max_dust_fraction))
    wlg = H.LogWavelengthGrid(0.1e-6, 10e-6, num_wavelengths)
    instr = H.FullInstrument(instrumentName="i60", distance=10e6 * pc, inclination=60 * DEG,
                             fieldOfViewX=40000 * pc, numPixelsX=num_pixels, fieldOfViewY=40000 * pc,
                             numPixelsY=num_pixels, recordComponents=True, recordStatistics=record_statistics)
    return H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr],
                                  numPackets=num_packets, minWavelength=0.1e-6, maxWavelength=10e-6,
                                  defaultWavelengthGrid=wlg, storeRadiationField=False, numDensitySamples=20,
                                  seed=seed)


def cfg4(num_packets=2e5, max_level=6, max_dust_fraction=2e-4, seed=0, min_level=3, num_sed_wavelengths=50,
         max_secondary_iterations=5, record_statistics=False):
    """dust emission with secondary-emission iterations; 1e4 Lsun 10 000 K point source in an r^-2 dust
    shell with tau_Z(0.55 um) = 20, octree, RF grid 40 bins 0.1-1000 um, emission grid 60 bins 1-1000 um,
    SEDInstrument with components.  (BASELINE.json configs[3] scales the tree to ~1e6 cells: max_level=8,
    max_dust_fraction~3e-6.)"""
    pc = H.PC
    mix = H.MeanListDustMix([0.05e-6, 0.55e-6, 10e-6, 100e-6, 2000e-6], [5000.0, 1000.0, 100.0, 5.0, 0.01],
                            [0.4, 0.6, 0.2, 0.01, 0.0001], [0.6, 0.5, 0.05, 0.0, 0.0])
    medium = H.GeometricMedium(H.ShellGeometry(0.01 * pc, 1.0 * pc, 2.0), mix, opticalDepth=20.0, wavelength=0.55e-6)
    grid = H.PolicyTreeSpatialGrid(-pc, pc, -pc, pc, -pc, pc, H.DensityTreePolicy(min_level, max_level, max_dust_fraction))
    src = H.PointSource((0.0, 0.0, 0.0), H.BlackBodySED(10000.0), luminosity=1e4 * H.LSUN)
    instr = H.SEDInstrument(instrumentName="sed", distance=1e6 * pc, inclination=60 * DEG, recordComponents=True,
                            recordStatistics=record_statistics)
    return H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr], numPackets=num_packets,
                                  minWavelength=0.1e-6, maxWavelength=20e-6,
                                  defaultWavelengthGrid=H.LogWavelengthGrid(0.1e-6, 1000e-6, num_sed_wavelengths),
                                  storeRadiationField=True, radiationFieldWLG=H.LogWavelengthGrid(0.1e-6, 1000e-6, 40),
                                  dustEmissionWLG=H.LogWavelengthGrid(1e-6, 1000e-6, 60), iterateSecondaryEmission=True,
                                  minSecondaryIterations=1, maxSecondaryIterations=max_secondary_iterations,
                                  numDensitySamples=20, seed=seed)


def cfg5(sites, num_packets=2e5, seed=0, num_pixels=64, density=None, volumes=None, record_statistics=True):
    """Voronoi grid on imported SPH particle positions (policy ImportedSites), exponential-disk source,
    one wavelength.  `sites` (m) are the particle positions; the cell densities are given (imported from the reference's
    ParticleMedium sampling) or default to a smooth disk evaluated at the sites."""
    pc = H.PC
    disk = H.ExpDiskGeometry(3000 * pc, 300 * pc, 0.0, 15000 * pc, 2000 * pc)
    src = H.GeometricSource(disk, H.BlackBodySED(6000.0), luminosity=1e10 * H.LSUN)
    mix = H.MeanListDustMix([0.1e-6, 1e-6], [1000.0, 1000.0], [0.6, 0.6], [0.5, 0.5])
    medium = H.GeometricMedium(H.ExpDiskGeometry(3000 * pc, 250 * pc, 0.0, 15000 * pc, 1900 * pc), mix, opticalDepth=1.0,
                               wavelength=0.55e-6)
    grid = H.VoronoiMeshSpatialGrid(-16000 * pc, 16000 * pc, -16000 * pc, 16000 * pc, -2000 * pc, 2000 * pc, sites,
                                    volumes=volumes)
    instr = H.FullInstrument(instrumentName="i60", distance=10e6 * pc, inclination=60 * DEG, fieldOfViewX=32000 * pc,
                             numPixelsX=num_pixels, fieldOfViewY=32000 * pc, numPixelsY=num_pixels, recordComponents=True,
                             recordStatistics=record_statistics)
    sim = H.MonteCarloSimulation(sources=[src], medium=medium, grid=grid, instruments=[instr], numPackets=num_packets,
                                 oligoWavelengths=[0.55e-6], storeRadiationField=False, numDensitySamples=1, seed=seed)
    sim.density = density
    return sim
