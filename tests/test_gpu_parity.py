"""Parity of the CUDA engine with the CPU oracle on the same seeded inputs, through the C ABI (B200 only)."""
import numpy as np
import pytest

from skirt9_b200 import abi
from skirt9_b200 import host as H
from tests import models
from tests.oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu


def run_both(sim, engine_lib):
    sim.setup()
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(gpu)
    sim.run(cpu)
    return gpu, cpu


def test_cartesian_cfg1(engine_lib):
    sim = models.small_cartesian(num_packets=30000, record_statistics=True)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["packets"] == 30000


def test_octree_cfg2_small(engine_lib):
    sim = models.small_octree(num_packets=30000, record_statistics=True)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_octree_deeper_levels(engine_lib):
    sim = models.small_octree(num_packets=10000, max_level=8, max_dust_fraction=2e-5, min_level=2)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert sim.grid.num_cells > 50000


def test_two_sources_three_instruments_forced(engine_lib):
    sim = models.two_sources_three_instruments(num_packets=20000, force=True)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_two_sources_three_instruments_nonforced(engine_lib):
    sim = models.two_sources_three_instruments(num_packets=20000, force=False)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_empty_medium_and_ragged_count(engine_lib):
    """Zero density everywhere: no scattering, direct == transparent; count not a multiple of the chunk size."""
    sim = models.small_cartesian(num_packets=1237)
    sim.setup()
    sim.density[:] = 0.0
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(gpu)
    sim.run(cpu)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["scatterings"] == 0
    np.testing.assert_allclose(gpu.read_sed(0, abi.SK_COMP_TRANSPARENT), gpu.read_sed(0, abi.SK_COMP_PRIMARY_DIRECT),
                               rtol=1e-12)


def test_history_sharding_matches_single_run(engine_lib):
    sim = models.small_octree(num_packets=8000)
    sim.setup()
    a = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    b = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(a)
    b.prepare_primary(8000)
    for first, count in ((0, 1), (1, 2999), (3000, 5000)):
        b.run_segment(first, count, True, True, False, 0)
    models.compare_engines(sim, a, b, rtol=1e-11)


def test_dust_emission_iterations_match_oracle(engine_lib):
    """DustEmission mode: primary emission, secondary-emission iterations to convergence, final secondary emission."""
    import copy
    sim = models.small_dust_emission(num_packets=20000).setup()
    simc = copy.copy(sim)
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = simc.configure(OracleEngine(simc.config_struct()))
    sim.run(gpu)
    simc.run(cpu)
    assert len(sim.convergence) == len(simc.convergence) >= 1
    for a, b in zip(sim.convergence, simc.convergence):
        assert a["converged"] == b["converged"]
        for key in ("dust_luminosity", "absorbed_primary", "absorbed_secondary"):
            assert a[key] == pytest.approx(b[key], rel=1e-9), key
    assert sim.dust_luminosity == pytest.approx(simc.dust_luminosity, rel=1e-9)
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)
    sec = gpu.read_sed(0, abi.SK_COMP_SECONDARY_DIRECT)
    assert sec.sum() > 0


def test_dust_emission_on_cartesian_grid(engine_lib):
    import copy
    sim = models.small_dust_emission(num_packets=8000)
    pc = 3.08567758e16
    from skirt9_b200 import host as H
    sim.grid = H.CartesianSpatialGrid(-pc, pc, -pc, pc, -pc, pc, 12, 10, 14)
    sim.setup()
    simc = copy.copy(sim)
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = simc.configure(OracleEngine(simc.config_struct()))
    sim.run(gpu)
    simc.run(cpu)
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)


def test_voronoi_grid_matches_oracle(engine_lib):
    sim = models.small_voronoi(num_packets=20000)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["forward_segments"] > 20000


def test_dust_emission_on_voronoi_grid(engine_lib):
    """Secondary emission launched from Voronoi cells: rejection sampling in the cell's enclosing box
    (VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989), with iterations."""
    sim = models.small_voronoi_dust_emission(num_packets=8000)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)
    assert sim.sed_flux_density(gpu, 0, abi.SK_COMP_SECONDARY_DIRECT).sum() > 0
    # without the cell extents the engine refuses instead of guessing
    sim2 = models.small_voronoi_dust_emission(num_packets=1000).setup()
    sim2.grid.cell_extents = None
    with pytest.raises(abi.SkError):
        sim2.configure(abi.Engine(sim2.config_struct(device=0), lib=engine_lib))


def test_voronoi_grid_with_radiation_field_nonforced_and_outside_source(engine_lib):
    """Voronoi grid with a stored radiation field, and a source that reaches beyond the grid (paths enter from outside)."""
    from skirt9_b200 import host as H
    sim = models.small_voronoi(num_packets=10000)
    pc = H.PC
    sim.storeRadiationField = True
    sim.sources[0].geometry = H.ExpDiskGeometry(6000 * pc, 900 * pc, 0.0, 25000 * pc, 4000 * pc)  # beyond the +-16/2 kpc box
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    sim2 = models.small_voronoi(num_packets=10000)
    sim2.forceScattering = False
    gpu, cpu = run_both(sim2, engine_lib)
    models.compare_engines(sim2, gpu, cpu)


def test_tabulated_sed_ring_source_and_forward_peaked_scattering(engine_lib):
    sim = models.tabulated_sed_ring_source_high_g(num_packets=20000)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED_LEVEL + 2).sum() > 0


def test_small_bank_refills_slots(engine_lib, monkeypatch):
    """A bank far smaller than the number of histories: slots are reused many times; results do not change."""
    sim = models.small_octree(num_packets=20000).setup()
    a = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(a)
    monkeypatch.setenv("SK_BANK", "1024")
    b = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(b)
    models.compare_engines(sim, a, b, rtol=1e-10)


def test_full_size_octree_properties(engine_lib):
    """BASELINE.json configs[1] at full grid size (~9.3e5 cells) with 2e6 packets: size-independent properties --
    transparent SED equals the analytic L_nu/(4 pi d^2) per bin up to wavelength-sampling noise, components are
    consistent, counters obey the path identities, and the result does not depend on how histories are sharded."""
    from skirt9_b200 import configs
    sim = configs.cfg2(num_packets=2e6).setup()
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(e)
    c = e.counters()
    assert c["packets"] == 2000000
    assert c["forward_paths"] == c["peel_paths"] == c["detections"] == c["scatterings"] + c["packets"]
    assert 30 < c["forward_segments"] / c["forward_paths"] < 80     # the reference measures 49 on its 933 059-cell tree
    tr = e.read_sed(0, abi.SK_COMP_TRANSPARENT)
    g = sim.defaultWavelengthGrid
    sed = sim.sources[0].sed
    want = np.array([np.trapz(sed.specific_luminosity(np.linspace(a, b, 400)), np.linspace(a, b, 400))
                     for a, b in zip(g.borderv[:-1], g.borderv[1:])]) * sim.sources[0].luminosity
    inside = (g.borderv[:-1] >= sim.source_range[0]) & (g.borderv[1:] <= sim.source_range[1])
    np.testing.assert_allclose(tr[inside], want[inside], rtol=0.02)
    di, sc = e.read_sed(0, abi.SK_COMP_PRIMARY_DIRECT), e.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED)
    assert np.all(di <= tr * (1 + 1e-12)) and np.all(sc >= 0)
    ifu = e.read_ifu(0, abi.SK_COMP_PRIMARY_DIRECT)
    assert ifu.sum(axis=1) == pytest.approx(di, rel=1e-3)            # the frame covers (almost) the whole model


def test_engine_rejects_unsupported_and_bad_calls(engine_lib):
    e = abi.Engine(abi.SkConfig(0, 1, 0, 0.5, 1e4, 0, 0), lib=engine_lib)
    with pytest.raises(abi.SkError) as ei:
        e.set_medium(np.ones(8))
    assert ei.value.code == abi.SK_ERR_STATE
    e.set_grid_cartesian([0, 1, 2], [0, 1, 2], [0, 1, 2])
    with pytest.raises(abi.SkError) as ei:
        e.set_medium(np.ones(7))
    assert ei.value.code == abi.SK_ERR_INVALID
    with pytest.raises(abi.SkError) as ei:
        e.run_segment(0, 10)
    assert ei.value.code == abi.SK_ERR_STATE
    with pytest.raises(abi.SkError) as ei:
        abi.Engine(abi.SkConfig(0, 1, 0, 0.5, 1e4, 99, 0), lib=engine_lib)
    assert ei.value.code == abi.SK_ERR_INVALID
    # malformed octree node lists are caught by the device-side checks of sk_engine_set_grid_octree
    box = [0, 0, 0, 1, 1, 1]
    leaves = [-1] * 8
    for bad in ([5] + leaves,                       # children beyond the end of the list
                [1] + leaves[:7] + [1] + leaves,    # a node pointing back at an earlier block: two parents / not after
                [1] + leaves + leaves):             # orphan nodes that no parent points to
        with pytest.raises(abi.SkError) as ei:
            e.set_grid_octree(box, bad)
        assert ei.value.code == abi.SK_ERR_INVALID, bad
    e.set_grid_octree(box, [1] + leaves[:7] + [9] + leaves)   # a valid two-level tree
    e.set_medium(np.ones(15))


def test_per_pixel_statistics_of_histories_with_hundreds_of_pixels(engine_lib, monkeypatch):
    """The per-history pixel list is exact beyond the SK_PIX_K = 32 entries of a bank slot (chunks chained from a pool):
    Sum w^k per pixel, k = 0..4, equals the oracle's, which keeps a list of any length like the reference's ContributionList."""
    sim = models.long_histories_many_pixels()
    gpu, cpu = run_both(sim, engine_lib)
    c = gpu.counters()
    assert c["scatterings"] / c["packets"] > 150          # long histories ...
    st = cpu.read_ifu_stats(0)
    assert st[0].sum() / c["packets"] > 100                # ... that reach > 100 distinct pixels each, on average
    assert c["pixel_overflows"] == 0
    models.compare_engines(sim, gpu, cpu)
    gpu.close()
    # a pool that is too small is reported, and leaves the first moments intact (only Sum w^k, k >= 2, sees the split lists)
    monkeypatch.setenv("SK_PIX_POOL", "8")
    small = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(small)
    assert small.counters()["pixel_overflows"] > 0
    a, b = small.read_ifu_stats(0), st
    np.testing.assert_allclose(a[1], b[1], rtol=1e-8, atol=1e-8 * b[1].max())
    assert a[0].sum() > b[0].sum()


@pytest.mark.parametrize("force", [True, False])
def test_explicit_absorption(engine_lib, force):
    """PhotonPacketOptions explicitAbsorption: interaction points in scattering optical depth, weight exp(-tau_abs)
    (MonteCarloSimulation.cpp:567-570, 727-731, 757-762), with forced (radiation field stored) and non-forced scattering,
    on an octree and on the ragged Cartesian grid with two sources and three instruments."""
    sim = models.two_sources_three_instruments(num_packets=20000, force=force)
    sim.explicitAbsorption = True
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    sim = models.small_octree(num_packets=20000, record_statistics=True)
    sim.explicitAbsorption = True
    sim.forceScattering = force
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_engine_limits_are_reported_as_unsupported(engine_lib):
    """More than 8 instruments or more than 8 individually recorded scattering levels are refused with SK_ERR_UNSUPPORTED
    (the shim lists the same limits in GpuLifeCycle::unsupportedReason), never silently truncated."""
    import copy
    sim = models.small_cartesian(num_packets=100)
    sim.setup()
    many = copy.copy(sim)
    many.instruments = [copy.copy(sim.instruments[0]) for _ in range(9)]
    with pytest.raises(abi.SkError) as err:
        many.configure(abi.Engine(many.config_struct(device=0), lib=engine_lib))
    assert err.value.code == abi.SK_ERR_UNSUPPORTED and "8 instruments" in str(err.value)
    deep = copy.copy(sim)
    deep.instruments = [copy.copy(sim.instruments[0])]
    deep.instruments[0].numScatteringLevels = 9
    with pytest.raises(abi.SkError) as err:
        deep.configure(abi.Engine(deep.config_struct(device=0), lib=engine_lib))
    assert err.value.code == abi.SK_ERR_UNSUPPORTED and "scattering levels" in str(err.value)


def test_interleaved_shares_add_up_to_the_single_run(engine_lib):
    """sk_engine_set_history_interleave: three engines that each run every third block of 256 histories of the same
    segment -- the sharding of the multi-GPU drivers -- together give the tallies of one engine running all of it."""
    sim = models.small_octree(num_packets=10000, record_statistics=True)
    sim.setup()
    one = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(one)
    parts = []
    for part in range(3):
        e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
        e.set_history_interleave(256, 3, part)
        sim.run(e)
        parts.append(e)
    assert sum(e.counters()["packets"] for e in parts) == one.counters()["packets"] == 10000
    assert [e.counters()["packets"] for e in parts] == [3344, 3328, 3328]
    for comp in (abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED):
        np.testing.assert_allclose(sum(e.read_sed(0, comp) for e in parts), one.read_sed(0, comp), rtol=1e-10)
        a, b = sum(e.read_ifu(0, comp) for e in parts), one.read_ifu(0, comp)
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12 * b.max())
    np.testing.assert_allclose(sum(e.read_sed_stats(0) for e in parts), one.read_sed_stats(0), rtol=1e-9)
    with pytest.raises(abi.SkError):
        parts[0].set_history_interleave(100, 3, 0)     # not a power of two


# ---------------------------------------------------------------- several medium components with their own mixes
def test_two_components_cartesian_forced_with_radiation_field(engine_lib):
    sim = models.with_second_component(models.two_sources_three_instruments(num_packets=20000, force=True), "shell")
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["scatterings"] > 20000
    # the second component matters: the same model without it gives other tallies
    ref = models.two_sources_three_instruments(num_packets=20000, force=True)
    one, _ = run_both(ref, engine_lib)
    assert not np.allclose(one.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED), gpu.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED), rtol=1e-3)


def test_two_components_cartesian_nonforced(engine_lib):
    sim = models.with_second_component(models.two_sources_three_instruments(num_packets=20000, force=False), "shell")
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_two_components_octree(engine_lib):
    sim = models.with_second_component(models.small_octree(num_packets=30000, record_statistics=True), "disk")
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert sim.density.shape[0] == 2 and (sim.density[1] > 0).any()


def test_three_components_voronoi(engine_lib):
    from skirt9_b200 import host as H
    sim = models.with_second_component(models.small_voronoi(num_packets=20000), "disk")
    third = H.MeanListDustMix([0.1e-6, 1e-6], [500.0, 700.0], [0.2, 0.9], [0.0, 0.8])
    sim.extraMedia.append(H.GeometricMedium(H.ExpDiskGeometry(5000 * H.PC, 400 * H.PC, 0.0, 15000 * H.PC, 1900 * H.PC), third,
                                            opticalDepth=0.3, wavelength=0.55e-6))
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_two_components_dust_emission_iterations(engine_lib):
    import copy
    sim = models.with_second_component(models.small_dust_emission(num_packets=20000), "shell").setup()
    simc = copy.copy(sim)
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = simc.configure(OracleEngine(simc.config_struct()))
    sim.run(gpu)
    simc.run(cpu)
    assert len(sim.convergence) == len(simc.convergence) >= 1
    for a, b in zip(sim.convergence, simc.convergence):
        for key in ("dust_luminosity", "absorbed_primary", "absorbed_secondary"):
            assert a[key] == pytest.approx(b[key], rel=1e-9), key
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)
    assert gpu.read_sed(0, abi.SK_COMP_SECONDARY_DIRECT).sum() > 0


def test_component_limits_are_reported(engine_lib):
    """More than SK_MAX_MEDIA components, mixes on different wavelength grids, a mix count that does not match the medium
    state: reported, never run."""
    sim = models.with_second_component(models.small_cartesian(num_packets=1000), "shell").setup()
    e = abi.Engine(sim.config_struct(device=0), lib=engine_lib)
    sim.grid.configure(e)
    with pytest.raises(abi.SkError):
        e.set_media(np.zeros((5, sim.grid.num_cells)), sim.volume)
    mixes = [(md.mix.lambda_border, md.mix.sigma_abs, md.mix.sigma_sca, md.mix.asymmpar, md.mix.mu) for md in sim.media]
    shifted = (mixes[1][0] * 1.01,) + mixes[1][1:]
    with pytest.raises(abi.SkError):
        e.set_dustmixes([mixes[0], shifted])
    e2 = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    e2.set_dustmix(*mixes[0])          # one mix for two components
    e2.prepare_primary(1000)
    with pytest.raises(abi.SkError):
        e2.run_segment(0, 1000, True, True, False, 0)


@pytest.mark.parametrize("force", [True, False])
def test_two_components_with_explicit_absorption(engine_lib, force):
    """Explicit absorption with several mixes (MediumSystem.cpp:937-955, 1112-1150): the walks are in scattering optical depth,
    the absorption optical depth -- whose ratio to it now varies from cell to cell -- is accumulated next to it and interpolated
    at the interaction point; the radiation field sees the sum of the two."""
    sim = models.with_second_component(models.two_sources_three_instruments(num_packets=20000, force=force), "shell")
    sim.explicitAbsorption = True
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["scatterings"] > 10000
    # against the same model with the albedo weighting: other weights, so other tallies
    ref = models.with_second_component(models.two_sources_three_instruments(num_packets=20000, force=force), "shell")
    one, _ = run_both(ref, engine_lib)
    assert not np.allclose(one.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED), gpu.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED), rtol=1e-3)


def test_three_components_octree_with_explicit_absorption(engine_lib):
    from skirt9_b200 import host as H
    sim = models.with_second_component(models.small_octree(num_packets=20000, record_statistics=True), "disk")
    third = H.MeanListDustMix([0.1e-6, 0.55e-6, 10e-6], [800.0, 900.0, 100.0], [0.0, 0.9, 0.5], [0.0, 0.7, 0.2])   # no scattering in the UV
    sim.extraMedia.append(H.GeometricMedium(H.ExpDiskGeometry(6000 * H.PC, 500 * H.PC, 0.0, 20000 * H.PC, 2000 * H.PC), third,
                                            opticalDepth=0.3, wavelength=0.55e-6))
    sim.explicitAbsorption = True
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


# ---------------------------------------------------------------- kinematics (sk_engine_set_velocities, moving sources)
@pytest.mark.parametrize("force", [True, False])
def test_kinematics_cartesian_two_sources(engine_lib, force):
    """Moving point and shell sources in an expanding medium on the ragged Cartesian grid: per-cell perceived wavelengths for
    the optical depths and (forced) the radiation field, Doppler shifts at launch, scattering and peel-off."""
    sim = models.with_kinematics(models.two_sources_three_instruments(num_packets=20000, force=force))
    if force:
        sim.radiationFieldWLG = H.LogWavelengthGrid(0.15e-6, 8e-6, 120)   # bins of 3 %: the shifts cross them
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["scatterings"] > 0


def test_kinematics_velocity_table_at_rest_equals_the_static_engine(engine_lib):
    """A velocity table of zeros sends the run through the kernels with per-cell look-ups; the tallies must be those of the
    single-medium kernels."""
    sim = models.small_octree(num_packets=20000, record_statistics=True)
    sim.setup()
    a = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    b = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    b.set_velocities(np.zeros((sim.grid.num_cells, 3)))
    sim.run(a)
    sim.run(b)
    models.compare_engines(sim, a, b, rtol=1e-11)


def test_kinematics_octree_moving_source_only(engine_lib):
    """Only the source moves (Configuration::hasMovingSources): the engine makes up the table of velocities at rest."""
    sim = models.with_kinematics(models.small_octree(num_packets=20000, record_statistics=True), media=False)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_kinematics_octree_ring_source_with_radiation_field(engine_lib):
    sim = models.with_kinematics(models.tabulated_sed_ring_source_high_g(num_packets=20000))
    sim.radiationFieldWLG = H.LogWavelengthGrid(0.15e-6, 8e-6, 80)
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_kinematics_dust_emission_iterations(engine_lib):
    """Dust cells emit with their bulk velocity (DustSecondarySource.cpp:562-580); the radiation field of the secondary
    segments is binned at the perceived wavelengths too."""
    sim = models.with_kinematics(models.small_dust_emission(num_packets=20000))
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert len(sim.convergence) >= 1


@pytest.mark.parametrize("force", [True, False])
def test_kinematics_two_components_with_explicit_absorption(engine_lib, force):
    sim = models.with_second_component(models.two_sources_three_instruments(num_packets=15000, force=force), kind="shell")
    sim = models.with_kinematics(sim)
    sim.explicitAbsorption = True
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_kinematics_single_component_with_explicit_absorption(engine_lib):
    sim = models.with_kinematics(models.small_octree(num_packets=15000))
    sim.explicitAbsorption = True
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_kinematics_voronoi(engine_lib):
    """Velocities per Voronoi cell handed over directly (the host mirror computes them for box-shaped cells only)."""
    sim = models.small_voronoi(num_packets=15000)
    sim.setup()
    rng = np.random.default_rng(4)
    sim.velocity = rng.normal(0.0, 3e6, size=(sim.grid.num_cells, 3))
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    cpu = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(gpu)
    sim.run(cpu)
    models.compare_engines(sim, gpu, cpu)


def test_velocities_are_dropped_with_the_medium_state(engine_lib):
    sim = models.small_cartesian(num_packets=2000)
    sim.setup()
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    with pytest.raises(abi.SkError):
        e.set_velocities(np.zeros((sim.grid.num_cells + 1, 3)))
    e.set_velocities(np.zeros((sim.grid.num_cells, 3)))
    e.set_medium(sim.density, sim.volume)      # a new medium state is at rest
    ref = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(e)
    sim.run(ref)
    models.compare_engines(sim, e, ref, rtol=1e-12)


def test_dynamic_state_iterations_match_oracle(engine_lib):
    """Primary and merged iterations with a ClearDensityRecipe through the host mirror: the engine is handed new densities between
    segments (sk_engine_set_medium on a configured engine keeps links, tables, radiation field and detector arrays)."""
    sim = models.small_dynamic_state(num_packets=20000)
    sim.setup()
    initial = np.array(sim.density, copy=True)
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(gpu)
    cleared_gpu, its_gpu = np.array(sim.density == 0), [it["updated_cells"] for it in sim.primary_iterations]
    sim.density = initial.copy()     # (the recipe has changed the model's densities: the oracle starts from the same state)
    cpu = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(cpu)
    assert its_gpu == [it["updated_cells"] for it in sim.primary_iterations] and sum(its_gpu) > 0
    assert np.array_equal(cleared_gpu, sim.density == 0)
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)


def test_dust_emission_with_cmb_heating(engine_lib):
    """DustEmissionOptions::includeHeatingByCMB at redshift 6: the CMB source term in the energy balance of every cell."""
    sim = models.small_dust_emission(num_packets=20000)
    sim.sources[0].luminosity = 3.0 * H.LSUN      # faint source: the 19 K background sets most dust temperatures
    sim.includeHeatingByCMB, sim.cosmologyRedshift = True, 6.0
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)


def test_kinematics_octree_nonforced(engine_lib):
    """Moving source and medium on the octree without forced scattering: the walk to the interaction point with per-cell
    look-ups (MediumSystem.cpp:1042-1070)."""
    sim = models.with_kinematics(models.small_octree(num_packets=30000, record_statistics=True))
    sim.forceScattering = False
    gpu, cpu = run_both(sim, engine_lib)
    models.compare_engines(sim, gpu, cpu)
    assert gpu.counters()["scatterings"] > 0
