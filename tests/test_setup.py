"""octree construction by the density policy and medium-state sampling, on the engine's side.

CPU part (`-m "not gpu"`): the oracle's restatement of DensityTreePolicy::constructTree / needsSubdivide
(DensityTreePolicy.cpp:116-309) and of the cell loop of MediumSystem::setupSelfAfter (MediumSystem.cpp:286-330) against
the policy's own semantics, against the numpy host mirror (independent code, independent random numbers) and against the
tree the unmodified reference built for tests/golden/cfg2s (statistical agreement: the reference samples with its
Mersenne twisters).
GPU part (`-m gpu`): the CUDA kernels of skirt9_b200/csrc/sk_setup.cuh against the oracle on the same Philox draws --
the node list must be identical, densities equal to 1e-12 (libm ulps), volumes bit-exact -- and a whole simulation set
up on the device against the reference's fluxes.
"""
import math
import os

import numpy as np
import pytest

from skirt9_b200 import abi, configs
from skirt9_b200 import host as H
from tests.oracle_lib import OracleEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PC = H.PC


def cfg2s(device_setup, num_packets=20000, **kw):
    args = dict(num_packets=num_packets, seed=0, max_level=6, max_dust_fraction=1e-4, num_pixels=64, num_wavelengths=10,
                record_statistics=True)
    args.update(kw)
    sim = configs.cfg2(**args)
    sim.deviceSetup = device_setup
    return sim.setup()


def levels_of(first_child):
    lev = np.zeros(len(first_child), dtype=int)
    for l in np.nonzero(first_child >= 0)[0]:
        lev[first_child[l]:first_child[l] + 8] = lev[l] + 1
    return lev


def build(engine, sim):
    sim.configure(engine)
    sim.fetch_device_setup(engine)
    return getattr(sim.grid, "first_child", None), sim.density, sim.volume


# ----------------------------------------------------------------------------------------------------------------------
# oracle (CPU)
# ----------------------------------------------------------------------------------------------------------------------

def test_oracle_tree_obeys_the_policy():
    sim = cfg2s(True)
    e = OracleEngine(sim.config_struct())
    fc, dens, vol = build(e, sim)
    pol = sim.grid.policy
    lev = levels_of(fc)
    # parent-before-child breadth-first list: levels ascend, children are consecutive blocks of 8
    assert np.all(np.diff(lev) >= 0)
    kids = np.sort(fc[fc >= 0])
    assert np.array_equal(kids, 1 + 8 * np.arange(len(kids)))
    # DensityTreePolicy.cpp:119-120
    assert np.all(fc[lev < pol.minLevel] >= 0)
    assert lev.max() <= pol.maxLevel and np.all(fc[lev == pol.maxLevel] < 0)
    # the mass criterion (DensityTreePolicy.cpp:192-196): with the cell's sampled mean density (an independent set of
    # samples) leaves above the minimum level hold about <= maxDustFraction of the mass, divided nodes more
    boxes = sim.grid.boxes
    leaf = fc < 0
    frac_leaf = dens * vol / sim.medium.number
    inner = leaf & (lev > pol.minLevel) & (lev < pol.maxLevel)
    assert np.quantile(frac_leaf[inner[leaf]], 0.99) < 2.0 * pol.maxDustFraction
    assert abs(vol.sum() - np.prod(np.array(sim.grid.extent[3:]) - np.array(sim.grid.extent[:3]))) < 1e-9 * vol.sum()
    assert boxes.shape == (len(fc), 6)
    # all the dust of the (normalised) geometry inside the box is accounted for: Sum n V = number x (mass in box)
    assert (dens * vol).sum() == pytest.approx(sim.medium.number, rel=0.02)


def test_oracle_tree_statistically_like_host_mirror_and_reference():
    dev = cfg2s(True)
    e = OracleEngine(dev.config_struct())
    fc, dens, vol = build(e, dev)
    host = cfg2s(False)  # numpy restatement, numpy random numbers
    n_oracle, n_host = int((fc < 0).sum()), host.grid.num_cells
    assert abs(n_oracle - n_host) < 0.03 * n_host, (n_oracle, n_host)
    # the tree the unmodified reference built for the same ski (tests/golden/make_golden.py)
    g = np.load(os.path.join(GOLD, "cfg2s_ref.npz"))
    n_ref = len(g["cell_volume_pc3"])
    assert abs(n_oracle - n_ref) < 0.03 * n_ref, (n_oracle, n_ref)
    # same distribution of cells over the levels (cell volume <-> level)
    lv_ref = np.round(np.log2(g["cell_volume_pc3"].max() / g["cell_volume_pc3"]) / 3).astype(int)
    lv_own = np.round(np.log2(vol.max() / vol) / 3).astype(int)
    h_ref = np.bincount(lv_ref, minlength=8)
    h_own = np.bincount(lv_own, minlength=8)
    assert np.all(np.abs(h_ref - h_own) <= 0.15 * h_ref + 16), (h_ref, h_own)  # 20 samples per node: noisy at the threshold
    # dust mass on the grid: reference densities are per-cell means of 20 samples as well
    RHO = H.MSUN / PC ** 3
    m_ref = (g["mass_density_msun_pc3"] * RHO * g["cell_volume_pc3"] * PC ** 3).sum()
    m_own = (dens * vol).sum() * dev.medium.mix.MU
    assert m_own == pytest.approx(m_ref, rel=0.02)


def shell_with_all_criteria(device_setup):
    """cfg4's r^-2 shell with the optical-depth and density-dispersion criteria next to the mass fraction
    (DensityTreePolicy.cpp:199-210)."""
    sim = configs.cfg4(num_packets=1000, seed=2, max_level=6)
    pol = sim.grid.policy
    sim.grid.policy = H.DensityTreePolicy(pol.minLevel, pol.maxLevel, 2e-3, maxDustOpticalDepth=0.2, wavelength=0.55e-6,
                                          maxDustDensityDispersion=0.9)
    sim.deviceSetup = device_setup
    return sim.setup()


def test_oracle_tree_with_optical_depth_and_dispersion_criteria():
    dev = shell_with_all_criteria(True)
    e = OracleEngine(dev.config_struct())
    fc, dens, vol = build(e, dev)
    host = shell_with_all_criteria(False)   # numpy restatement with its own random numbers
    n_oracle, n_host = int((fc < 0).sum()), host.grid.num_cells
    assert abs(n_oracle - n_host) < 0.05 * n_host, (n_oracle, n_host)
    # each criterion matters: dropping it gives a smaller tree
    for drop in ("maxDustOpticalDepth", "maxDustDensityDispersion"):
        sim = shell_with_all_criteria(True)
        setattr(sim.grid.policy, drop, 0.0)
        e2 = OracleEngine(sim.config_struct())
        fc2, _, _ = build(e2, sim)
        assert (fc2 < 0).sum() < 0.985 * n_oracle, drop
    # the optical-depth criterion holds for the leaves below the maximum level (sampled mean density x diagonal)
    lev = levels_of(fc)
    leaf = fc < 0
    boxes = dev.grid.boxes[leaf]
    diag = np.linalg.norm(boxes[:, 3:] - boxes[:, :3], axis=1)
    tau = dev.grid.policy.dust_kappa([dev.medium]) * dens * dev.medium.mix.mu * diag
    inner = lev[leaf] < dev.grid.policy.maxLevel
    assert np.quantile(tau[inner], 0.98) < 2.0 * dev.grid.policy.maxDustOpticalDepth


def test_oracle_density_sampling_cartesian():
    sim = configs.cfg1(num_packets=1000, seed=0)
    sim.deviceSetup = True
    sim.numDensitySamples = 1
    sim.setup()
    e = OracleEngine(sim.config_struct())
    _, dens, vol = build(e, sim)
    boxes = sim.grid.cell_boxes()
    c = 0.5 * (boxes[:, :3] + boxes[:, 3:])
    np.testing.assert_allclose(dens, sim.medium.number_density(c[:, 0], c[:, 1], c[:, 2]), rtol=1e-13)
    np.testing.assert_allclose(vol, np.prod(boxes[:, 3:] - boxes[:, :3], axis=1), rtol=1e-15)
    # 100 samples per cell: unbiased mean of the density over the cell
    sim.numDensitySamples = 100
    e2 = OracleEngine(sim.config_struct())
    _, d100, _ = build(e2, sim)
    inside = np.linalg.norm(np.abs(c) + 0.5 * (boxes[:, 3:] - boxes[:, :3]), axis=1) < sim.medium.geometry.rmax
    np.testing.assert_allclose(d100[inside], dens[inside], rtol=1e-12)  # uniform sphere: constant inside
    assert (d100 * vol).sum() == pytest.approx(sim.medium.number, rel=0.01)


def test_unsupported_setup_is_refused():
    sim = cfg2s(True)
    e = OracleEngine(sim.config_struct())
    with pytest.raises(abi.SkError):
        e.sample_medium(sim.medium.density_geometry(), 20, 0)   # no grid yet
    with pytest.raises(abi.SkError):
        e.read_octree()


# ----------------------------------------------------------------------------------------------------------------------
# CUDA engine against the oracle and the reference
# ----------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("case", ["cfg2s", "cfg2_full", "cfg4_shell", "all_criteria"])
def test_device_tree_and_densities_equal_oracle(case):
    def make():
        if case == "cfg2s":
            return cfg2s(True)
        if case == "cfg2_full":
            sim = configs.cfg2(num_packets=1000, seed=0)   # 933 k cells, levels 3-9
            sim.deviceSetup = True
            return sim.setup()
        if case == "all_criteria":
            return shell_with_all_criteria(True)
        sim = configs.cfg4(num_packets=1000, seed=2, max_level=6)   # shell geometry (power law), dust emission setup
        sim.deviceSetup = True
        return sim.setup()
    a, b = make(), make()
    eo = OracleEngine(a.config_struct())
    eg = abi.Engine(b.config_struct())
    fo, do, vo = build(eo, a)
    fg, dg, vg = build(eg, b)
    assert len(fo) == len(fg)
    assert np.array_equal(fo, fg)
    assert np.array_equal(vo, vg)
    np.testing.assert_allclose(dg, do, rtol=1e-12, atol=0)
    if case == "cfg2_full":
        assert 0.9e6 < (fg < 0).sum() < 0.97e6


@pytest.mark.gpu
def test_device_setup_cartesian_equals_oracle():
    sims = []
    for _ in range(2):
        sim = configs.cfg1(num_packets=20000, seed=4)
        sim.deviceSetup = True
        sim.numDensitySamples = 50
        sims.append(sim.setup())
    eo, eg = OracleEngine(sims[0].config_struct()), abi.Engine(sims[1].config_struct())
    _, do, vo = build(eo, sims[0])
    _, dg, vg = build(eg, sims[1])
    assert np.array_equal(vo, vg)
    np.testing.assert_allclose(dg, do, rtol=1e-12)
    # and the life cycle runs on the state the device built: same tallies as the oracle on its own copy
    from tests.models import compare_engines
    sims[0].run(eo)
    sims[1].run(eg)
    compare_engines(sims[1], eo, eg)


@pytest.mark.gpu
def test_simulation_set_up_on_the_device_matches_reference_fluxes():
    """cfg2s end to end with the grid built and the densities sampled by the CUDA kernels: the SED agrees with the
    unmodified reference (2e7-packet fixture, its own tree and densities) within the Monte-Carlo error plus the
    grid-discretisation scatter between two independently sampled trees (measured below 1 %)."""
    n = 2_000_000
    sim = cfg2s(True, num_packets=n)
    e = abi.Engine(sim.config_struct())
    sim.configure(e)
    sim.run(e)
    g = np.load(os.path.join(GOLD, "cfg2s_hi_ref.npz"))
    sed = g["sed"]
    tot = sim.sed_flux_density(e, 0, abi.SK_COMP_TOTAL)
    tr = sim.sed_flux_density(e, 0, abi.SK_COMP_TRANSPARENT)
    st = e.read_sed_stats(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_own = np.sqrt(np.maximum(st[2] / (st[1] * st[1]) - 1.0 / np.maximum(st[0], 1), 0.0))
    assert np.all(np.abs(tr - sed[:, 2]) <= (4 * r_own + 1e-6) * sed[:, 2])
    assert np.all(np.abs(tot - sed[:, 1]) <= (4 * r_own + 0.01) * sed[:, 1]), (tot / sed[:, 1])


# ---------------------------------------------------------------- Voronoi tessellation (VoronoiMeshSnapshot::buildMesh, voro++)
def disk_sites(n, seed=99):
    rng = np.random.default_rng(seed)
    R = np.minimum(rng.gamma(2.0, 3000.0, size=n), 15000.0)
    phi = rng.uniform(0, 2 * np.pi, size=n)
    z = np.clip(rng.laplace(0.0, 300.0, size=n), -1900.0, 1900.0)
    s = np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1) * PC
    return s[np.argsort(s[:, 0], kind="stable")]


VOR_EXTENT = (-16000 * PC, -16000 * PC, -2000 * PC, 16000 * PC, 16000 * PC, 2000 * PC)


def test_oracle_voronoi_tessellation_is_a_partition_with_symmetric_faces():
    """The cells fill the domain box exactly, every face is shared by the two cells it separates, and volumes, enclosing
    boxes and neighbours agree with an independent construction (scipy's Qhull on the sites mirrored in the six walls)."""
    sites = disk_sites(2500)
    e = OracleEngine(configs.cfg1(num_packets=10).config_struct())
    entries = e.build_voronoi(VOR_EXTENT, sites)
    off, idx, vol, box = e.read_voronoi()
    n = len(sites)
    assert off[0] == 0 and off[-1] == entries == len(idx) and np.all(np.diff(off) >= 4)
    ext = np.asarray(VOR_EXTENT)
    assert vol.sum() == pytest.approx(np.prod(ext[3:] - ext[:3]), rel=1e-13)
    sets = [set(idx[off[m]:off[m + 1]].tolist()) for m in range(n)]
    assert all(m in sets[j] for m in range(n) for j in sets[m] if j >= 0)
    assert all(len(sets[m]) == off[m + 1] - off[m] for m in range(n))        # no neighbour listed twice
    grid = H.VoronoiMeshSpatialGrid(*(VOR_EXTENT[i] for i in (0, 3, 1, 4, 2, 5)), sites)
    grid.setup([], 1, None)
    grid.compute_cell_geometry()
    np.testing.assert_allclose(vol, grid.volumes, rtol=1e-11)
    np.testing.assert_allclose(box, grid.cell_extents, rtol=0, atol=1e-12 * (ext[3] - ext[0]))
    # every face neighbour is a Delaunay neighbour; Delaunay neighbours without a face have it outside the domain box
    for m in range(n):
        delaunay = set(x for x in grid.nbr_index[grid.nbr_offset[m]:grid.nbr_offset[m + 1]].tolist() if x >= 0)
        assert set(x for x in sets[m] if x >= 0) <= delaunay
    # the walls a cell lists are those its enclosing box touches
    for w in range(6):
        axis, upper = w >> 1, w & 1
        touches = box[:, axis + 3] >= ext[axis + 3] * (1 - 1e-14) if upper else box[:, axis] <= ext[axis] * (1 - 1e-14)
        listed = np.array([-(w + 1) in sets[m] for m in range(n)])
        assert np.array_equal(touches, listed), w


def test_oracle_voronoi_volumes_equal_the_reference():
    """All cell volumes of the cfg5s fixture -- written by the unmodified reference, whose tessellation is voro++'s -- to the
    10 digits of its text output.  The reference read the sites with the 9 digits of the particle file."""
    g = np.load(os.path.join(GOLD, "cfg5s_ref.npz"))
    pos = np.array([[float("%.8e" % v) for v in row[:3]] for row in g["particles"]]) * PC
    pos = pos[np.argsort(pos[:, 0], kind="stable")]          # cells in order of increasing x, VoronoiMeshSnapshot.cpp:507-508
    e = OracleEngine(configs.cfg1(num_packets=10).config_struct())
    e.build_voronoi(VOR_EXTENT, pos)
    vol = e.read_voronoi()[2]
    np.testing.assert_allclose(vol, g["cell_volume_pc3"] * PC ** 3, rtol=1e-9)


def test_voronoi_builder_refuses_what_it_cannot_do():
    e = OracleEngine(configs.cfg1(num_packets=10).config_struct())
    x = (np.arange(4) + 0.5) / 4 * 2 - 1
    lattice = np.stack(np.meshgrid(x, x, x, indexing="ij"), axis=-1).reshape(-1, 3) * PC      # four planes through every vertex
    with pytest.raises(abi.SkError):
        e.build_voronoi((-PC, -PC, -PC, PC, PC, PC), lattice)
    with pytest.raises(abi.SkError):
        e.build_voronoi((-PC, -PC, -PC, PC, PC, PC), np.array([[0.1, 0.2, 0.3], [0.1, 0.2, 0.3]]) * PC)   # coinciding sites
    with pytest.raises(abi.SkError):
        e.build_voronoi((-PC, -PC, -PC, PC, PC, PC), np.array([[0.1, 0.2, 1.5]]) * PC)                      # outside the domain


def test_life_cycle_on_the_built_tessellation_equals_the_delaunay_lists():
    """The face neighbours are all a walk needs: the same run on the built lists and on the (larger) Delaunay neighbour lists."""
    from tests import models
    runs = []
    for device_setup in (True, False):
        sim = models.small_voronoi(num_packets=4000)
        sim.deviceSetup = device_setup
        sim.setup()
        e = sim.configure(OracleEngine(sim.config_struct()))
        sim.run(e)
        runs.append((sim, e))
    assert runs[0][0].grid.nbr_offset[-1] < runs[1][0].grid.nbr_offset[-1]
    models.compare_engines(runs[0][0], runs[0][1], runs[1][1])


@pytest.mark.gpu
def test_device_voronoi_tessellation_equals_oracle():
    """sk_voronoi_build_kernel against the oracle: the same lists in the same order, volumes and boxes bit for bit."""
    for n in (300, 20000):
        sites = disk_sites(n, seed=n)
        cfg = configs.cfg1(num_packets=10).config_struct(device=0)
        gpu, cpu = abi.Engine(cfg), OracleEngine(cfg)
        assert gpu.build_voronoi(VOR_EXTENT, sites) == cpu.build_voronoi(VOR_EXTENT, sites)
        a, b = gpu.read_voronoi(), cpu.read_voronoi()
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[3], b[3])
    x = (np.arange(4) + 0.5) / 4 * 2 - 1
    lattice = np.stack(np.meshgrid(x, x, x, indexing="ij"), axis=-1).reshape(-1, 3) * PC
    with pytest.raises(abi.SkError):
        abi.Engine(cfg).build_voronoi((-PC, -PC, -PC, PC, PC, PC), lattice)


@pytest.mark.gpu
def test_voronoi_simulation_set_up_on_the_device_equals_oracle():
    from tests import models
    sim = models.small_voronoi_dust_emission(num_packets=6000)
    sim.deviceSetup = True
    sim.setup()
    gpu = sim.configure(abi.Engine(sim.config_struct(device=0)))
    import copy
    simc = copy.copy(sim)
    cpu = simc.configure(OracleEngine(simc.config_struct()))
    sim.run(gpu)
    simc.run(cpu)
    models.compare_engines(sim, gpu, cpu, rtol=1e-8)
    assert sim.sed_flux_density(gpu, 0, abi.SK_COMP_SECONDARY_DIRECT).sum() > 0


# ---------------------------------------------------------------- ParticleMedium: smoothed-particle density per cell
MSUN = H.MSUN


def sph_fixture(name):
    """The particle table of a fixture as the reference read it (9 digits in the text file), in SI units."""
    g = np.load(os.path.join(GOLD, name + "_ref.npz"))
    part = np.array([[float("%.8e" % v) for v in row] for row in g["particles"]])
    part[:, :4] *= PC
    part[:, 4] *= MSUN
    return g, part


def cartesian_24_20_10(engine):
    engine.set_grid_cartesian(np.linspace(-16000 * PC, 16000 * PC, 25), np.linspace(-16000 * PC, 16000 * PC, 21),
                              np.linspace(-2000 * PC, 2000 * PC, 11))
    return 24 * 20 * 10


def test_oracle_particle_density_equals_the_reference():
    """tests/golden/ski/cfg13p.ski: ParticleMedium on a Cartesian grid with numDensitySamples = 1 -- the reference's
    ParticleSnapshot::density at every cell centre, deterministic; to the 10 digits of its text output, and the same set of
    empty cells."""
    g, part = sph_fixture("cfg13p")
    e = OracleEngine(configs.cfg1(num_packets=10).config_struct())
    nc = cartesian_24_20_10(e)
    e.sample_medium_particles(part, 1.0, 1, nc)
    dens, vol = e.read_medium()
    ref = g["mass_density_msun_pc3"] * MSUN / PC ** 3
    assert np.array_equal(dens > 0, ref > 0) and (ref > 0).sum() > 1500
    np.testing.assert_allclose(dens[ref > 0], ref[ref > 0], rtol=1e-9)
    np.testing.assert_allclose(vol, g["cell_volume_pc3"] * PC ** 3, rtol=1e-9)


def test_oracle_particle_density_in_voronoi_cells_like_the_reference():
    """cfg5s: ten random positions per Voronoi cell.  The reference draws them from its Mersenne twister, so the agreement is
    statistical: total mass, and the cell values around the reference's within the scatter of ten samples."""
    g, part = sph_fixture("cfg5s")
    sites = part[:, :3][np.argsort(part[:, 0], kind="stable")]
    e = OracleEngine(configs.cfg1(num_packets=10).config_struct())
    e.build_voronoi(VOR_EXTENT, sites)
    e.sample_medium_particles(part, 1.0, 10, len(sites))
    dens, vol = e.read_medium()
    ref = g["mass_density_msun_pc3"] * MSUN / PC ** 3
    refvol = g["cell_volume_pc3"] * PC ** 3
    np.testing.assert_allclose(vol, refvol, rtol=1e-9)
    assert (dens * vol).sum() == pytest.approx((ref * refvol).sum(), rel=0.02)
    ok = (ref > 0) & (dens > 0)
    assert ok.mean() > 0.99
    assert np.median(dens[ok] / ref[ok]) == pytest.approx(1.0, abs=0.02)
    assert np.corrcoef(np.log(dens[ok]), np.log(ref[ok]))[0, 1] > 0.85
    with pytest.raises(abi.SkError):
        e.sample_medium_particles(part, 1.0, 1, len(sites))      # the centroid of a Voronoi cell is not available


@pytest.mark.gpu
def test_device_particle_density_equals_oracle():
    """sk_sample_particles_kernel against the oracle on a Cartesian grid (cell centres and random positions), an octree and a
    Voronoi grid built on the device: the same positions, the same particles in the same order."""
    g, part = sph_fixture("cfg13p")
    cfg = configs.cfg1(num_packets=10, seed=4).config_struct(device=0)
    for kind in ("cartesian1", "cartesian7", "octree", "voronoi"):
        gpu, cpu = abi.Engine(cfg), OracleEngine(cfg)
        if kind.startswith("cartesian"):
            nc = [cartesian_24_20_10(x) for x in (gpu, cpu)][0]
            ns = int(kind[-1])
        elif kind == "octree":
            sim = cfg2s(False, num_packets=10)
            for x in (gpu, cpu):
                sim.grid.configure(x)
            nc, ns = sim.grid.num_cells, 3
        else:
            sites = part[:, :3][np.argsort(part[:, 0], kind="stable")][::3]
            for x in (gpu, cpu):
                x.build_voronoi(VOR_EXTENT, sites)
            nc, ns = len(sites), 4
        for x in (gpu, cpu):
            x.sample_medium_particles(part, 2.5, ns, nc)
        (da, va), (db, vb) = gpu.read_medium(), cpu.read_medium()
        assert (db > 0).sum() > 0.3 * nc, kind
        np.testing.assert_allclose(da, db, rtol=1e-13, atol=0, err_msg=kind)
        np.testing.assert_array_equal(va, vb, err_msg=kind)


@pytest.mark.gpu
def test_particle_medium_on_a_device_built_voronoi_grid_runs_like_the_reference():
    """cfg5s end to end with nothing but the particle table: sites sorted and tessellated, densities sampled and the life cycle run
    on the device; fluxes against the reference's (its set-up differs in the random positions only)."""
    from tests.test_golden_reference import check_cfg5s
    g, part = sph_fixture("cfg5s")
    n = 2000000
    sim = configs.cfg5(part[:, :3], num_packets=n, seed=0)
    sim.medium = H.ParticleMedium(part, sim.medium.mix)
    sim.numDensitySamples = 10
    sim.deviceSetup = True
    sim.setup()
    e = sim.configure(abi.Engine(sim.config_struct(device=0)))
    sim.run(e)
    check_cfg5s(sim, e, g, n)
