"""The drop-in itself: shim/_build/skirt_b200 = the UNMODIFIED reference (its .ski reader, setup code, probes, FITS and text
writers, linked from the reference's object files) with the photon life cycle handed to the GPU engine by the C++ shim
shim/GpuLifeCycle.cpp.  The very .ski files the golden fixtures were produced from are run unchanged, and the output
FILES are compared with what the reference's own CPU life cycle wrote for them (tests/golden/*.npz).  `-t 1` makes the
reference's setup (tree construction, density sampling) identical to the fixture run, so the densities agree bit for bit
and only the random streams of the life cycle differ."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

from tests.skirt_files import read_columns, read_fits_cube

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "shim", "_build", "skirt_b200")
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


def run_ski(name, tmp_path, packets, devices=None, host_setup=True):
    """host_setup=True passes --host-setup: the reference's own (CPU, -t 1) tree construction, so that grid and densities
    are bit for bit those of the fixture run; False lets the drop-in construct the octree on the GPU."""
    if not os.path.exists(EXE):
        pytest.fail("shim/_build/skirt_b200 has not been built (make -C shim, where /root/reference exists)")
    text = open(os.path.join(GOLD, "ski", name + ".ski")).read()
    text = re.sub(r'numPackets="[^"]*"', 'numPackets="%g"' % packets, text, count=1)
    ski = tmp_path / (name + ".ski")
    ski.write_text(text)
    extra = (["-g", devices] if devices else []) + (["--host-setup"] if host_setup else [])
    subprocess.check_call([EXE, "-t", "1", "-b", "-o", str(tmp_path)] + extra + [str(ski)], stdout=subprocess.DEVNULL)
    log = (tmp_path / (name + "_log.txt")).read_text()
    assert "GPU life cycle:" in log, log[-2000:]
    return log


def rel_error(stats_row):
    n, w1, w2 = stats_row[0], stats_row[1], stats_row[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.maximum(w2 / (w1 * w1) - 1.0 / np.maximum(n, 1), 0.0))


def sed_columns_within_statistics(sed, stats, ref, ref_stats, cols, nsigma=4.0, nsigma_secondary=5.0, secondary_per_bin=True):
    """|F - F_ref| <= nsigma sqrt(R^2 + R_ref^2) max(F_ref, F_ref_total) for the bins of the given columns whose error
    estimate is reliable on both sides by the reference's own rule (R < 0.1, VOV < 0.1; tests/mcstats.py), R from both
    sides' Sum w^k statistics; the columns with dust emission (5-7 and the total) get nsigma_secondary,
    because the statistics of the last segment do not contain the noise of the radiation field behind the dust
    temperatures -- against a fixture with few packets (2e5 per segment) that noise dominates, and those columns are then
    compared through their sums only (secondary_per_bin=False).  The sums over the bins agree within the quadrature sum of
    the bins' errors."""
    from tests import mcstats
    own, rst = stats[:, 1:].T, ref_stats[:, 1:].T
    ok = mcstats.reliable(own) & mcstats.reliable(rst)
    assert ok.sum() >= 0.8 * len(ok)
    sigma = np.hypot(mcstats.rel_error(own), mcstats.rel_error(rst))
    for col in cols:
        ns = nsigma if col in (2, 3, 4) else nsigma_secondary
        scale = np.maximum(ref[:, col], ref[:, 1]) * sigma
        z = np.abs(sed[:, col] - ref[:, col])[ok] / np.maximum(scale, 1e-300)[ok]
        if col in (2, 3, 4) or secondary_per_bin:
            assert np.all(z <= ns), (col, int(np.argmax(z)), float(z.max()))
        assert abs(sed[ok, col].sum() - ref[ok, col].sum()) <= ns * np.sqrt((scale[ok] ** 2).sum()), col


def test_cfg1_ski_runs_unchanged(tmp_path):
    g = np.load(os.path.join(GOLD, "cfg1_ref.npz"))
    hi = np.load(os.path.join(GOLD, "cfg1_hi_ref.npz"))
    n = 4e6
    log = run_ski("cfg1", tmp_path, n)
    assert re.search(r"Finished primary emission in [0-9.]+ s", log)
    sed = read_columns(tmp_path / "cfg1_i60_sed.dat")[0]
    stats = read_columns(tmp_path / "cfg1_i60_sedstats.dat")[0]
    cells = read_columns(tmp_path / "cfg1_cells_cellprops.dat")
    # same setup as the fixture run: the reference's own density sampling, bit for bit
    np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])
    assert stats[1] == n
    ref = g["sed"][0]
    assert sed[2] == pytest.approx(ref[2], rel=1e-8)     # transparent flux: noise free
    assert sed[3] == pytest.approx(ref[3], rel=1e-8)     # direct flux: same densities, same optical depth
    for r in (g, hi):
        tol = 4.0 * math.hypot(rel_error(r["sedstats"][0, 1:]), rel_error(stats[1:]))
        assert abs(sed[1] - r["sed"][0, 1]) <= tol * r["sed"][0, 1]
        assert abs(sed[4] - r["sed"][0, 4]) <= tol * r["sed"][0, 1]
    for comp in ("transparent", "primarydirect"):
        frame, cards = read_fits_cube(tmp_path / f"cfg1_i60_{comp}.fits")
        np.testing.assert_allclose(frame, g["frame_" + comp], rtol=3e-6)
        assert cards["BUNIT"] == "MJy/sr"
    total, _ = read_fits_cube(tmp_path / "cfg1_i60_total.fits")
    assert total.sum() == pytest.approx(hi["frame_total_sum"].sum(), rel=2e-3)
    # per-pixel statistics frames written by the reference from the engine's Sum w^k tallies: histories per pixel
    s0, _ = read_fits_cube(tmp_path / "cfg1_i60_stats0.fits")
    assert s0.astype(float).sum() / (n / 1e6) == pytest.approx(float(g["frame_stats0"].sum()), rel=0.005)
    assert s0.max() > 0 and (tmp_path / "cfg1_i60_stats4.fits").exists()
    # the radiation field probe of the reference, fed from the engine's tally
    J = read_columns(tmp_path / "cfg1_rf_J.dat")[:, 1]
    assert J.sum() == pytest.approx(g["J_nu"][:, 0].sum(), rel=0.004)


def test_cfg2s_ski_octree_runs_unchanged(tmp_path):
    g = np.load(os.path.join(GOLD, "cfg2s_ref.npz"))
    hi = np.load(os.path.join(GOLD, "cfg2s_hi_ref.npz"))
    n = 4e6
    run_ski("cfg2s", tmp_path, n)
    sed = read_columns(tmp_path / "cfg2s_i60_sed.dat")
    stats = read_columns(tmp_path / "cfg2s_i60_sedstats.dat")
    cells = read_columns(tmp_path / "cfg2s_cells_cellprops.dat")
    np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])   # same tree, same densities
    np.testing.assert_allclose(sed[:, 0], g["sed"][:, 0], rtol=1e-9)
    r_own = rel_error(stats[:, 1:].T)
    for r in (g, hi):
        tol = 4.0 * np.hypot(rel_error(r["sedstats"][:, 1:].T), r_own)
        for col in (1, 2, 3, 4):
            bound = tol * np.maximum(r["sed"][:, col], r["sed"][:, 1])
            assert np.all(np.abs(sed[:, col] - r["sed"][:, col]) <= bound), col
    total, _ = read_fits_cube(tmp_path / "cfg2s_i60_total.fits")
    assert total.astype(float).sum() == pytest.approx(hi["frame_total_sum"].sum(), rel=3e-3)


def test_cfg2s_ski_with_the_octree_constructed_on_the_gpu(tmp_path):
    """SURVEY.md 8f row f2 in the drop-in: the DensityTreePolicy of the ski file builds its octree with
    sk_engine_build_octree (Philox samples instead of the reference's Mersenne twister): the tree agrees statistically with
    the fixture's (cell count, dust mass on the grid) and the SED within the Monte-Carlo error plus 1 %."""
    g = np.load(os.path.join(GOLD, "cfg2s_ref.npz"))
    hi = np.load(os.path.join(GOLD, "cfg2s_hi_ref.npz"))
    log = run_ski("cfg2s", tmp_path, 4e6, host_setup=False)
    assert "The spatial tree will be constructed on the GPU" in log and "GPU tree construction:" in log
    cells = read_columns(tmp_path / "cfg2s_cells_cellprops.dat")
    nref = len(g["mass_density_msun_pc3"])
    assert abs(len(cells) - nref) <= 0.03 * nref
    mass, mass_ref = (cells[:, 6] * cells[:, 4]).sum(), (g["mass_density_msun_pc3"] * g["cell_volume_pc3"]).sum()
    assert mass == pytest.approx(mass_ref, rel=0.01)
    # the topology probe of the reference walks the node objects the shim rebuilt from the engine's node list
    topo = [int(t) for t in (tmp_path / "cfg2s_topo_treetop.dat").read_text().split("\n") if t and not t.startswith("#")]
    assert topo.count(0) == len(cells)
    sed = read_columns(tmp_path / "cfg2s_i60_sed.dat")
    stats = read_columns(tmp_path / "cfg2s_i60_sedstats.dat")
    tol = 4.0 * np.hypot(rel_error(hi["sedstats"][:, 1:].T), rel_error(stats[:, 1:].T)) + 0.01
    for col in (1, 2, 3, 4):
        bound = tol * np.maximum(hi["sed"][:, col], hi["sed"][:, 1])
        assert np.all(np.abs(sed[:, col] - hi["sed"][:, col]) <= bound), col


def test_cfg9e_ski_explicit_absorption_runs_unchanged(tmp_path):
    """cfg1 with explicitAbsorption="true" (MonteCarloSimulation.cpp:567-570, 727-731) against the reference's run of it."""
    g = np.load(os.path.join(GOLD, "cfg9e_ref.npz"))
    n = 4e6
    run_ski("cfg9e", tmp_path, n)
    sed = read_columns(tmp_path / "cfg9e_i60_sed.dat")[0]
    stats = read_columns(tmp_path / "cfg9e_i60_sedstats.dat")[0]
    ref = g["sed"][0]
    assert sed[2] == pytest.approx(ref[2], rel=1e-8)     # transparent flux: noise free
    assert sed[3] == pytest.approx(ref[3], rel=1e-8)     # direct flux: same densities, same optical depth
    tol = 4.0 * math.hypot(rel_error(g["sedstats"][0, 1:]), rel_error(stats[1:]))
    assert abs(sed[1] - ref[1]) <= tol * ref[1]
    assert abs(sed[4] - ref[4]) <= tol * ref[1]
    J = read_columns(tmp_path / "cfg9e_rf_J.dat")[:, 1]
    assert J.sum() == pytest.approx(g["J_nu"][:, 0].sum(), rel=0.004)


def test_cfg8z_ski_observer_frame_redshift_runs_unchanged(tmp_path):
    """cfg2s seen from redshift 0.5 (FlatUniverseCosmology, instrument distance 0): the engine bins the packets at
    lambda (1 + z) (FluxRecorder.cpp:309-310), the reference's writer calibrates with the luminosity distance."""
    g = np.load(os.path.join(GOLD, "cfg8z_ref.npz"))
    log = run_ski("cfg8z", tmp_path, 4e6)
    assert "redshift" not in re.findall(r"outside the GPU life cycle \(([^)]*)\)", log)
    sed = read_columns(tmp_path / "cfg8z_i60_sed.dat")
    stats = read_columns(tmp_path / "cfg8z_i60_sedstats.dat")
    head = open(tmp_path / "cfg8z_i60_sed.dat").readline()
    assert "redshift 0.5" in head
    np.testing.assert_allclose(sed[:, 0], g["sed"][:, 0], rtol=1e-9)
    tol = 4.0 * np.hypot(rel_error(g["sedstats"][:, 1:].T), rel_error(stats[:, 1:].T))
    for col in (1, 2, 3, 4):
        bound = tol * np.maximum(g["sed"][:, col], g["sed"][:, 1])
        assert np.all(np.abs(sed[:, col] - g["sed"][:, col]) <= bound), col
    total, _ = read_fits_cube(tmp_path / "cfg8z_i60_total.fits")
    assert total.astype(float).sum() == pytest.approx(g["frame_total_sum"].sum(), rel=4e-3)


def test_cfg10d_ski_two_media_sharing_one_mix_runs_unchanged(tmp_path):
    """Two dust media with the same material mix (ring + exponential disk): the reference's several-media configuration
    (MediumSystem.cpp:874-885); the shim hands the engine the summed density.  Once with the reference's own set-up (same tree
    and densities as the fixture), once with the tree built on the device from BOTH geometries."""
    g = np.load(os.path.join(GOLD, "cfg10d_ref.npz"))
    for host_setup in (True, False):
        d = tmp_path / ("host" if host_setup else "device")
        d.mkdir()
        log = run_ski("cfg10d", d, 4e6, host_setup=host_setup)
        assert "GPU life cycle:" in log and "outside the GPU life cycle" not in log
        cells = read_columns(d / "cfg10d_cells_cellprops.dat")
        if host_setup:
            np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])   # both media, summed by the probe
        else:
            assert "GPU tree construction" in log
            assert len(cells) == pytest.approx(len(g["mass_density_msun_pc3"]), rel=0.03)
            mass = lambda c, rho: float((c * rho).sum())
            assert mass(cells[:, 4], cells[:, 6]) == pytest.approx(mass(g["cell_volume_pc3"], g["mass_density_msun_pc3"]), rel=0.01)
        sed = read_columns(d / "cfg10d_i60_sed.dat")
        stats = read_columns(d / "cfg10d_i60_sedstats.dat")
        tol = 4.0 * np.hypot(rel_error(g["sedstats"][:, 1:].T), rel_error(stats[:, 1:].T)) + (0.0 if host_setup else 0.01)
        for col in (1, 2, 3, 4):
            bound = tol * np.maximum(g["sed"][:, col], g["sed"][:, 1])
            assert np.all(np.abs(sed[:, col] - g["sed"][:, col]) <= bound), (host_setup, col)
        total, _ = read_fits_cube(d / "cfg10d_i60_total.fits")
        assert total.astype(float).sum() == pytest.approx(g["frame_total_sum"].sum(), rel=4e-3 if host_setup else 1.2e-2)


def test_cfg11m_ski_two_media_with_different_mixes_runs_unchanged(tmp_path):
    """Two dust media with DIFFERENT material mixes: the shim hands the engine one component per medium
    (sk_engine_set_media / sk_engine_set_dustmixes) and the several-media life cycle runs on the device."""
    g = np.load(os.path.join(GOLD, "cfg11m_ref.npz"))
    log = run_ski("cfg11m", tmp_path, 4e6)
    assert "GPU life cycle:" in log and "outside the GPU life cycle" not in log
    for h in range(2):   # same inputs as the fixture: the reference's own set-up, -t 1, seed 0
        rho = read_columns(tmp_path / ("cfg11m_dns_%d_rho.dat" % h))[:, 1]
        np.testing.assert_array_equal(rho, g["component_mass_density_msun_pc3"][h])
    sed = read_columns(tmp_path / "cfg11m_i60_sed.dat")
    stats = read_columns(tmp_path / "cfg11m_i60_sedstats.dat")
    tol = 4.0 * np.hypot(rel_error(g["sedstats"][:, 1:].T), rel_error(stats[:, 1:].T))
    for col in (1, 2, 3, 4):
        bound = tol * np.maximum(g["sed"][:, col], g["sed"][:, 1])
        assert np.all(np.abs(sed[:, col] - g["sed"][:, col]) <= bound), col
    total, _ = read_fits_cube(tmp_path / "cfg11m_i60_total.fits")
    assert total.astype(float).sum() == pytest.approx(g["frame_total_sum"].sum(), rel=4e-3)


def test_cfg14em_ski_two_mixes_with_explicit_absorption_runs_unchanged(tmp_path):
    g = np.load(os.path.join(GOLD, "cfg14em_ref.npz"))
    log = run_ski("cfg14em", tmp_path, 4e6)
    assert "GPU life cycle:" in log and "outside the GPU life cycle" not in log
    sed = read_columns(tmp_path / "cfg14em_i60_sed.dat")
    stats = read_columns(tmp_path / "cfg14em_i60_sedstats.dat")
    tol = 4.0 * np.hypot(rel_error(g["sedstats"][:, 1:].T), rel_error(stats[:, 1:].T))
    for col in (1, 2, 3, 4):
        bound = tol * np.maximum(g["sed"][:, col], g["sed"][:, 1])
        assert np.all(np.abs(sed[:, col] - g["sed"][:, col]) <= bound), col
    total, _ = read_fits_cube(tmp_path / "cfg14em_i60_total.fits")
    assert total.astype(float).sum() == pytest.approx(g["frame_total_sum"].sum(), rel=4e-3)


def test_cfg12me_ski_dust_emission_from_two_mixes_runs_unchanged(tmp_path):
    g = np.load(os.path.join(GOLD, "cfg12me_ref.npz"))
    log = run_ski("cfg12me", tmp_path, 2e6)
    prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
    sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
    conv = re.search(r"Convergence reached after (\d+) iterations", log)
    assert conv and int(conv.group(1)) == int(g["converged_after"])
    np.testing.assert_allclose(prim, g["absorbed_primary_lsun"], rtol=0.004)
    np.testing.assert_allclose(sec, g["absorbed_secondary_lsun"], rtol=0.02)
    sed = read_columns(tmp_path / "cfg12me_sed_sed.dat")
    stats = read_columns(tmp_path / "cfg12me_sed_sedstats.dat")
    hi = np.load(os.path.join(GOLD, "cfg12me_hi_ref.npz"))
    sed_columns_within_statistics(sed, stats, g["sed"], g["sedstats"], range(1, 8), secondary_per_bin=False)
    sed_columns_within_statistics(sed, stats, hi["sed"], hi["sedstats"], range(1, 8))


def test_cfg4s_ski_dust_emission_runs_unchanged(tmp_path):
    g = np.load(os.path.join(GOLD, "cfg4s_ref.npz"))
    n = 2e6
    log = run_ski("cfg4s", tmp_path, n)
    prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
    sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
    conv = re.search(r"Convergence reached after (\d+) iterations", log)
    assert conv and int(conv.group(1)) == int(g["converged_after"])
    np.testing.assert_allclose(prim, g["absorbed_primary_lsun"], rtol=0.004)
    np.testing.assert_allclose(sec, g["absorbed_secondary_lsun"], rtol=0.02)
    sed = read_columns(tmp_path / "cfg4s_sed_sed.dat")
    stats = read_columns(tmp_path / "cfg4s_sed_sedstats.dat")
    hi = np.load(os.path.join(GOLD, "cfg4s_hi_ref.npz"))   # the same ski with 2e6 packets per segment
    sed_columns_within_statistics(sed, stats, g["sed"], g["sedstats"], range(1, 8), secondary_per_bin=False)
    sed_columns_within_statistics(sed, stats, hi["sed"], hi["sedstats"], range(1, 8))
    # the reference's TemperatureProbe evaluated on the radiation field the engine handed back
    T = read_columns(tmp_path / "cfg4s_temp_dust_T.dat")[:, 1]
    ok = g["temperature"] > 0
    assert np.median(np.abs(T[ok] / g["temperature"][ok] - 1)) < 0.02
    assert np.array_equal(T > 0, g["temperature"] > 0) or np.mean((T > 0) != ok) < 0.01


@pytest.mark.parametrize("name, n", [("cfg15k", 2e6), ("cfg18ke", 1e6), ("cfg19ks", 2e6)])
def test_cfg15k_ski_kinematics_runs_unchanged(tmp_path, name, n):
    """Moving source, expanding dust shell, dust emission: the ski of the cfg15k fixture through the drop-in.  The reference's
    set-up samples densities and bulk velocities; the shim hands MediumState::bulkVelocity(m) and the source's velocity to the
    engine.  The two fine-grid SED instruments on opposite lines of sight resolve the Doppler shifts.  cfg18ke: the same with a
    second, rotating component of another mix and explicit absorption.  cfg19ks: a ring source with a rotation velocity field
    (GeometricSource + CylindricalVectorField) in dust at rest."""
    from tests import mcstats
    g = np.load(os.path.join(GOLD, name + "_ref.npz"))
    log = run_ski(name, tmp_path, n)
    assert ("Including support for kinematics" in log) == (name != "cfg19ks")   # (cfg19ks: only the source moves)
    lum = float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1))
    assert lum == pytest.approx(float(g["dust_luminosity_lsun"]), rel=0.005)
    for ins in ("fwd", "bwd", "sed"):
        sed = read_columns(tmp_path / ("%s_%s_sed.dat" % (name, ins)))
        own = read_columns(tmp_path / ("%s_%s_sedstats.dat" % (name, ins)))[:, 1:].T
        ref_sed, ref = g["sed_" + ins], g["sedstats_" + ins][:, 1:].T
        # (N of FluxRecorder.hpp:50-63: the packets launched in the two peel-off segments)
        ok = mcstats.reliable(own, launched=2 * n) & mcstats.reliable(ref, launched=2 * float(g["num_packets"]))
        assert ok.sum() >= 0.7 * len(ok)
        sigma = np.hypot(mcstats.rel_error(own, 2 * n), mcstats.rel_error(ref, 2 * float(g["num_packets"])))
        for col in range(1, 8):
            scale = np.maximum(ref_sed[:, col], ref_sed[:, 1])
            z = (np.abs(sed[:, col] - ref_sed[:, col]) / np.maximum(sigma * scale, 1e-300))[ok]
            assert np.all(z <= (4.5 if col <= 4 else 6.5)), (ins, col, int(np.argmax(z)), float(z.max()))
    # the shifted feature: 0.5-0.51 micron at rest, the source moves at 0.02 c along the line of sight of "fwd"
    # (cfg19ks: a ring that rotates at 0.02 c, the feature is a symmetric rotation profile about its rest wavelength)
    for ins, factor in ((("fwd", 1.0), ("bwd", 1.0)) if name == "cfg19ks" else (("fwd", 1 - 0.02001), ("bwd", 1 + 0.02001))):
        sed = read_columns(tmp_path / ("%s_%s_sed.dat" % (name, ins)))
        f = np.where(sed[:, 2] > 0.5 * sed[:, 2].max(), sed[:, 2] / sed[:, 0] ** 2, 0.0)   # F_nu -> F_lambda
        assert float((f * sed[:, 0]).sum() / f.sum()) == pytest.approx(0.505 * factor, rel=3e-3)


def test_cfg20kn_ski_kinematics_octree_nonforced_runs_unchanged(tmp_path):
    """A rotating disk (source velocity field, moving dust ring) on the octree without forced scattering: the ski of the cfg20kn
    fixture through the drop-in, with the reference's own (CPU, -t 1) tree so that grid and densities are the fixture's."""
    from tests import mcstats
    g = np.load(os.path.join(GOLD, "cfg20kn_ref.npz"))
    n = 4e6
    log = run_ski("cfg20kn", tmp_path, n)
    assert "Including support for kinematics" in log and "no forced scattering" in log
    for ins in ("edge", "i60"):
        sed = read_columns(tmp_path / ("cfg20kn_%s_sed.dat" % ins))
        own = read_columns(tmp_path / ("cfg20kn_%s_sedstats.dat" % ins))[:, 1:].T
        ref_sed, ref = g["sed_" + ins], g["sedstats_" + ins][:, 1:].T
        ok = mcstats.reliable(own, launched=n) & mcstats.reliable(ref, launched=float(g["num_packets"]))
        assert ok.sum() >= 4
        sigma = np.hypot(mcstats.rel_error(own, n), mcstats.rel_error(ref, float(g["num_packets"])))
        for col in range(1, 5):
            scale = np.maximum(ref_sed[:, col], ref_sed[:, 1])
            z = (np.abs(sed[:, col] - ref_sed[:, col]) / np.maximum(sigma * scale, 1e-300))[ok]
            assert np.all(z <= 4.5), (ins, col, int(np.argmax(z)), float(z.max()))


def test_cfg16d_ski_dynamic_state_iterations_run_unchanged(tmp_path):
    """Primary emission iterations and merged iterations with a ClearDensityRecipe (the call sites MonteCarloSimulation.cpp:314,
    451, 474): the segments run on the engine, the reference's own recipe code updates the medium state on the host from the
    radiation field handed back, and the new densities go to the engine.  Same packets as the fixture run: the sequence of
    cleared cells is the reference's up to the noise of the cells at the threshold."""
    g = np.load(os.path.join(GOLD, "cfg16d_ref.npz"))
    log = run_ski("cfg16d", tmp_path, 4e5)
    updated = [int(x) for x in re.findall(r"Updated cells: (\d+) out of", log)]
    nprim = len(re.findall(r"Finished primary emission iteration", log))
    nmerged = len(re.findall(r"Finished merged primary and secondary emission iteration", log))
    assert 2 <= nprim <= 5 and 1 <= nmerged <= 4
    assert len(updated) == nprim + nmerged and updated[0] > 0
    assert abs(sum(updated) - int(g["updated_cells"].sum())) <= 12
    lum = [float(x) for x in re.findall(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log)]
    prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
    sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
    assert lum[-1] == pytest.approx(g["dust_luminosity_lsun"][-1], rel=0.03)
    assert prim[-1] == pytest.approx(g["absorbed_primary_lsun"][-1], rel=0.03)
    assert sec[-1] == pytest.approx(g["absorbed_secondary_lsun"][-1], rel=0.08)
    # the medium state the reference's probe writes after the run: the cavity
    cells = read_columns(tmp_path / "cfg16d_cells_cellprops.dat")
    own = (cells[:, 6] == 0) & (g["initial_mass_density_msun_pc3"] > 0)
    ref = (g["final_mass_density_msun_pc3"] == 0) & (g["initial_mass_density_msun_pc3"] > 0)
    assert own.sum() == sum(updated)
    assert np.count_nonzero(own & ref) >= 0.8 * min(own.sum(), ref.sum())
    sed = read_columns(tmp_path / "cfg16d_sed_sed.dat")
    for col in (1, 2, 3, 4, 5):
        assert sed[:, col].sum() == pytest.approx(g["sed"][:, col].sum(), rel=0.01 if col == 2 else 0.05), col


def test_cfg17c_ski_cmb_heating_runs_unchanged(tmp_path):
    """Dust heated by the CMB at redshift 6, observer-frame instrument: the shim hands the calculator's CMB source term to the
    engine (sk_secondary_t::rf_cmb); the reference's TemperatureProbe then shows the 19 K floor."""
    from tests import mcstats
    g = np.load(os.path.join(GOLD, "cfg17c_ref.npz"))
    n = 1e6
    log = run_ski("cfg17c", tmp_path, n)
    lum = float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1))
    assert lum == pytest.approx(float(g["dust_luminosity_lsun"]), rel=0.01)
    sed = read_columns(tmp_path / "cfg17c_sed_sed.dat")
    own = read_columns(tmp_path / "cfg17c_sed_sedstats.dat")[:, 1:].T
    ref = g["sedstats"][:, 1:].T
    ok = mcstats.reliable(own, launched=2 * n) & mcstats.reliable(ref, launched=2 * float(g["num_packets"]))
    sigma = np.hypot(mcstats.rel_error(own, 2 * n), mcstats.rel_error(ref, 2 * float(g["num_packets"])))
    for col in range(1, 8):
        scale = np.maximum(g["sed"][:, col], g["sed"][:, 1])
        z = (np.abs(sed[:, col] - g["sed"][:, col]) / np.maximum(sigma * scale, 1e-300))[ok]
        assert np.all(z <= (4.5 if col <= 4 else 6.5)), (col, int(np.argmax(z)), float(z.max()))
    T = read_columns(tmp_path / "cfg17c_temp_dust_T.dat")[:, 1]
    filled = g["temperature"] > 0
    assert np.array_equal(T > 0, filled)
    assert T[filled].min() > 19.0 and np.median(np.abs(T[filled] / g["temperature"][filled] - 1)) < 0.01


def test_cfg5s_ski_voronoi_particles_runs_unchanged(tmp_path):
    """ParticleMedium import + VoronoiMeshSpatialGrid (voro++ tessellation, SPH kernel density sampling) all done by the
    reference's own setup code; only the life cycle runs on the GPU."""
    g = np.load(os.path.join(GOLD, "cfg5s_ref.npz"))
    rows = g["particles"]
    head = "# column 1: x (pc)\n# column 2: y (pc)\n# column 3: z (pc)\n# column 4: h (pc)\n# column 5: M (Msun)\n"
    (tmp_path / "sph.txt").write_text(head + "\n".join(" ".join("%.8e" % v for v in row) for row in rows) + "\n")
    n = 2e6
    text = open(os.path.join(GOLD, "ski", "cfg5s.ski")).read()
    text = re.sub(r'numPackets="[^"]*"', 'numPackets="%g"' % n, text, count=1)
    ski = tmp_path / "cfg5s.ski"
    ski.write_text(text)
    subprocess.check_call([EXE, "-t", "1", "-b", "-i", str(tmp_path), "-o", str(tmp_path), str(ski)],
                          stdout=subprocess.DEVNULL)
    log = (tmp_path / "cfg5s_log.txt").read_text()
    assert "GPU life cycle:" in log, log[-2000:]
    cells = read_columns(tmp_path / "cfg5s_cells_cellprops.dat")
    np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])   # same tessellation, same densities
    sed = read_columns(tmp_path / "cfg5s_i60_sed.dat")[0]
    stats = read_columns(tmp_path / "cfg5s_i60_sedstats.dat")[0]
    ref = g["sed"][0]
    assert sed[2] == pytest.approx(ref[2], rel=1e-8)
    tol = 4.0 * math.hypot(rel_error(g["sedstats"][0, 1:]), rel_error(stats[1:]))
    for col in (1, 3, 4):
        assert abs(sed[col] - ref[col]) <= tol * ref[1], (col, sed[col], ref[col], tol)


def test_cfg6m_ski_many_features_runs_unchanged(tmp_path):
    """Two sources (ring + ListSED, point + blackbody), forward-peaked dust, stored radiation field, three instruments on two
    lines of sight with aperture, scattering levels, statistics, roll, offsets and an instrument-specific wavelength grid."""
    g = np.load(os.path.join(GOLD, "cfg6m_ref.npz"))
    n = 4e6
    run_ski("cfg6m", tmp_path, n)
    cells = read_columns(tmp_path / "cfg6m_cells_cellprops.dat")
    np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])
    stats = read_columns(tmp_path / "cfg6m_sed_sedstats.dat")
    sed = read_columns(tmp_path / "cfg6m_sed_sed.dat")
    ref = g["sed_sed"]
    tol = 4.5 * np.hypot(rel_error(g["sed_stats"][:, 1:].T), rel_error(stats[:, 1:].T))
    # columns: total, transparent, direct, scattered, (3 secondary), 1-times and 2-times scattered
    for col in (1, 2, 3, 4, 8, 9):
        bound = tol * np.maximum(ref[:, col], ref[:, 1])
        assert np.all(np.abs(sed[:, col] - ref[:, col]) <= bound), (col, sed[:, col] / ref[:, col] - 1, tol)
    assert np.all(sed[:, 5:8] == 0)
    # the FullInstrument has its own 4-bin wavelength grid
    fsed, fref = read_columns(tmp_path / "cfg6m_full_sed.dat"), g["full_sed"]
    assert fsed.shape == fref.shape == (4, 8)
    np.testing.assert_allclose(fsed[:, 0], fref[:, 0], rtol=1e-9)
    np.testing.assert_allclose(fsed[:, 1:4], fref[:, 1:4], rtol=0.03)
    np.testing.assert_allclose(fsed[:, 4], fref[:, 4], rtol=0.08)   # the scattered flux inside the small 1 pc frame: few packets
    # frames: the rolled off-centre frame (total only) and the components of the full instrument
    frame, _ = read_fits_cube(tmp_path / "cfg6m_frame_total.fits")
    assert frame.shape == g["frame_total"].shape == (7, 14, 18)
    a, b = frame.astype(float).sum(axis=0), g["frame_total"].astype(float).sum(axis=0)
    ok = b > 0.05 * b.max()
    np.testing.assert_allclose(a[ok], b[ok], rtol=0.08)
    assert a.sum() == pytest.approx(b.sum(), rel=0.005)
    for comp in ("total", "transparent", "primarydirect", "primaryscattered"):
        x, _ = read_fits_cube(tmp_path / f"cfg6m_full_{comp}.fits")
        y = g["full_" + comp]
        assert x.shape == y.shape == (4, 8, 8)
        assert x.astype(float).sum() == pytest.approx(y.astype(float).sum(), rel=0.01), comp
    # the reference's radiation field probe on the engine's tally
    J = read_columns(tmp_path / "cfg6m_rf_J.dat")[:, 1:]
    np.testing.assert_allclose(J.sum(axis=0), g["J_nu"].astype(float).sum(axis=0), rtol=0.01)


def test_cfg7v_ski_voronoi_dust_emission_runs_unchanged(tmp_path):
    """Dust emission with iterations on a Voronoi grid (3000 sites drawn by the reference's set-up): the secondary packets
    are launched from Voronoi cells by rejection in the cells' enclosing boxes, which the shim reads from the reference's
    VoronoiMeshSnapshot::Cell objects."""
    g = np.load(os.path.join(GOLD, "cfg7v_ref.npz"))
    n = 2e6
    log = run_ski("cfg7v", tmp_path, n)
    cells = read_columns(tmp_path / "cfg7v_cells_cellprops.dat")
    np.testing.assert_array_equal(cells[:, 6], g["mass_density_msun_pc3"])   # same sites, same sampled densities
    prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
    sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
    conv = re.search(r"Convergence reached after (\d+) iterations", log)
    assert conv and int(conv.group(1)) == int(g["converged_after"])
    np.testing.assert_allclose(prim, g["absorbed_primary_lsun"], rtol=0.004)
    np.testing.assert_allclose(sec, g["absorbed_secondary_lsun"], rtol=0.02)
    sed = read_columns(tmp_path / "cfg7v_sed_sed.dat")
    stats = read_columns(tmp_path / "cfg7v_sed_sedstats.dat")
    hi = np.load(os.path.join(GOLD, "cfg7v_hi_ref.npz"))   # the same ski with 2e6 packets per segment
    sed_columns_within_statistics(sed, stats, g["sed"], g["sedstats"], range(1, 8), secondary_per_bin=False)
    sed_columns_within_statistics(sed, stats, hi["sed"], hi["sedstats"], range(1, 8))
    T = read_columns(tmp_path / "cfg7v_temp_dust_T.dat")[:, 1]
    ok = g["temperature"] > 0
    assert np.median(np.abs(T[ok] / g["temperature"][ok] - 1)) < 0.02


def _num_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs (skirt_b200 -g 0,1)")
def test_two_gpus_from_the_drop_in_binary(tmp_path):
    """skirt_b200 -g 0,1: one engine per device, contiguous history blocks, NCCL all-reduce of the radiation field after
    every segment and of the detector arrays before output, from C++ (shim/GpuLifeCycle.cpp)."""
    # primary emission only: octree, SED against the reference (same checks as the single-GPU test)
    g = np.load(os.path.join(GOLD, "cfg2s_ref.npz"))
    hi = np.load(os.path.join(GOLD, "cfg2s_hi_ref.npz"))
    d1 = tmp_path / "a"
    d1.mkdir()
    log = run_ski("cfg2s", d1, 4e6, devices="0,1")
    assert "2 devices, NCCL all-reduce" in log
    sed = read_columns(d1 / "cfg2s_i60_sed.dat")
    stats = read_columns(d1 / "cfg2s_i60_sedstats.dat")
    m = re.search(r"GPU life cycle: (\d+) packets", log)
    assert m and 3990000 <= int(m.group(1)) <= 4000000   # every history ran once, on one of the two devices
    r_own = rel_error(stats[:, 1:].T)
    for r in (g, hi):
        tol = 4.0 * np.hypot(rel_error(r["sedstats"][:, 1:].T), r_own)
        for col in (1, 2, 3, 4):
            bound = tol * np.maximum(r["sed"][:, col], r["sed"][:, 1])
            assert np.all(np.abs(sed[:, col] - r["sed"][:, col]) <= bound), col
    # dust emission with iterations: the radiation field is summed over the devices inside the iteration loop
    g4 = np.load(os.path.join(GOLD, "cfg4s_ref.npz"))
    d2 = tmp_path / "b"
    d2.mkdir()
    log = run_ski("cfg4s", d2, 2e6, devices="0,1")
    prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
    sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
    conv = re.search(r"Convergence reached after (\d+) iterations", log)
    assert conv and int(conv.group(1)) == int(g4["converged_after"])
    np.testing.assert_allclose(prim, g4["absorbed_primary_lsun"], rtol=0.004)
    np.testing.assert_allclose(sec, g4["absorbed_secondary_lsun"], rtol=0.02)
    sed4 = read_columns(d2 / "cfg4s_sed_sed.dat")
    for col in range(1, 8):
        assert sed4[:, col].sum() == pytest.approx(g4["sed"][:, col].sum(), rel=0.02)
