#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, built from
/root/reference by oracle/ref.mk) on the ski files under tests/golden/ski/ with `-t 1` (bit-reproducible).

The reference has no tests or golden vectors of its own, so these outputs ARE the pin of the oracle:
each fixture holds both the reference's INPUTS as the reference itself sampled them (per-cell densities from
SpatialCellPropertiesProbe, the octree topology from TreeSpatialGridTopologyProbe) and its OUTPUTS (SED columns,
Sigma w^k statistics, calibrated frames, per-cell mean intensity, convergence log values), so that the oracle and the
CUDA engine can be run on IDENTICAL inputs and compared within the Monte-Carlo tolerance the reference's own
statistics define.

    python tests/golden/make_golden.py [cfg1 cfg2s cfg4s cfg5s]

Only this script needs oracle/_ref; the tests read the committed .npz files.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SKIRT = os.path.join(ROOT, "oracle", "_ref", "release", "SKIRT", "main", "skirt")


sys.path.insert(0, ROOT)
from tests.skirt_files import read_columns, read_fits_cube  # noqa: E402


def run_reference(ski_name, workdir, extra_inputs=None, threads=1, packets=None):
    ski = os.path.join(HERE, "ski", ski_name + ".ski")
    if packets is not None:
        text = re.sub(r'numPackets="[^"]*"', 'numPackets="%g"' % packets, open(ski).read(), count=1)
        ski = os.path.join(workdir, ski_name + ".ski")
        open(ski, "w").write(text)
    for name, text in (extra_inputs or {}).items():
        open(os.path.join(workdir, name), "w").write(text)
    subprocess.check_call([SKIRT, "-t", str(threads), "-b", "-i", workdir, "-o", workdir, ski],
                          stdout=subprocess.DEVNULL)
    return open(os.path.join(workdir, ski_name + "_log.txt")).read()


def frames(workdir, prefix, names):
    out = {}
    for n in names:
        p = os.path.join(workdir, f"{prefix}_{n}.fits")
        if os.path.exists(p):
            out["frame_" + n] = read_fits_cube(p)[0]
    return out


def parse_topology(path):
    return np.array([int(t) for t in open(path).read().split("\n") if t and not t.startswith("#")], dtype=np.int8)


def make_cfg1():
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg1", d)
        sed = read_columns(os.path.join(d, "cfg1_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg1_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg1_cells_cellprops.dat"))
        rfJ = read_columns(os.path.join(d, "cfg1_rf_J.dat"))
        out = dict(sed=sed, sedstats=stats, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4].astype(np.float32),
                   cell_volume_pc3=cells[:, 4], J_nu=rfJ[:, 1:], num_packets=1e6,
                   seconds=float(re.search(r"Finished primary emission in ([0-9.]+) s", log).group(1)))
        out.update(frames(d, "cfg1_i60", ["total", "transparent", "primarydirect", "primaryscattered", "stats0",
                                          "stats1", "stats2"]))
    np.savez_compressed(os.path.join(HERE, "cfg1_ref.npz"), **out)
    print("cfg1:", {k: np.shape(v) for k, v in out.items()})


def make_cfg2s():
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg2s", d)
        sed = read_columns(os.path.join(d, "cfg2s_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg2s_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg2s_cells_cellprops.dat"))
        topo = parse_topology(os.path.join(d, "cfg2s_topo_treetop.dat"))
        out = dict(sed=sed, sedstats=stats, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4],
                   cell_volume_pc3=cells[:, 4], topology=topo, num_packets=1e6,
                   seconds=float(re.search(r"Finished primary emission in ([0-9.]+) s", log).group(1)))
        out.update(frames(d, "cfg2s_i60", ["total", "transparent", "primarydirect", "primaryscattered", "stats0",
                                           "stats1", "stats2"]))
    np.savez_compressed(os.path.join(HERE, "cfg2s_ref.npz"), **out)
    print("cfg2s:", {k: np.shape(v) for k, v in out.items()})


def make_cfg9e():
    """cfg1 with explicit absorption: same grid and densities (same seed, same set-up), 1e6 packets."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg9e", d)
        sed = read_columns(os.path.join(d, "cfg9e_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg9e_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg9e_cells_cellprops.dat"))
        base = np.load(os.path.join(HERE, "cfg1_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(cells[:, 6], base), "cfg9e must see the densities of cfg1"
        rfJ = read_columns(os.path.join(d, "cfg9e_rf_J.dat"))
        out = dict(sed=sed, sedstats=stats, J_nu=rfJ[:, 1:], num_packets=1e6)
        out.update(frames(d, "cfg9e_i60", ["total", "primaryscattered"]))
    np.savez_compressed(os.path.join(HERE, "cfg9e_ref.npz"), **out)
    print("cfg9e:", {k: np.shape(v) for k, v in out.items()}, sed)


def make_cfg8z():
    """cfg2s in the observer frame at redshift 0.5 (instrument distance 0): same tree and densities as cfg2s (same seed,
    same set-up), the packets binned at lambda (1 + z), the calibration with the luminosity distance."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg8z", d, packets=2e6)
        sed = read_columns(os.path.join(d, "cfg8z_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg8z_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg8z_cells_cellprops.dat"))
        base = np.load(os.path.join(HERE, "cfg2s_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(cells[:, 6], base), "cfg8z must see the tree and the densities of cfg2s"
        total = read_fits_cube(os.path.join(d, "cfg8z_i60_total.fits"))[0].astype(np.float64)
        head = open(os.path.join(d, "cfg8z_i60_sed.dat")).readline()
        dl = float(re.search(r"luminosity distance ([0-9.eE+-]+) Mpc", head).group(1))
        out = dict(sed=sed, sedstats=stats, frame_total_sum=total.sum(axis=0), num_packets=2e6, redshift=0.5,
                   luminosity_distance_mpc=dl)
    np.savez_compressed(os.path.join(HERE, "cfg8z_ref.npz"), **out)
    print("cfg8z:", {k: np.shape(v) for k, v in out.items()})


def make_cfg10d():
    """cfg2s with two dust media sharing one mix (ring + exponential disk): the reference's several-media code path; the fixture
    holds the tree and the TOTAL dust density the reference sampled.  2e6 packets."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg10d", d, packets=2e6)
        sed = read_columns(os.path.join(d, "cfg10d_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg10d_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg10d_cells_cellprops.dat"))
        topo = parse_topology(os.path.join(d, "cfg10d_topo_treetop.dat"))
        total = read_fits_cube(os.path.join(d, "cfg10d_i60_total.fits"))[0].astype(np.float64)
        out = dict(sed=sed, sedstats=stats, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4],
                   cell_volume_pc3=cells[:, 4], topology=topo, num_packets=2e6, frame_total_sum=total.sum(axis=0))
    np.savez_compressed(os.path.join(HERE, "cfg10d_ref.npz"), **out)
    print("cfg10d:", {k: np.shape(v) for k, v in out.items()})


def make_cfg11m():
    """Two dust media with DIFFERENT material mixes on the cfg2s octree (extinction only): the fixture holds the tree and the
    mass density of EACH component (DensityProbe, aggregation Component) as the reference sampled them.  2e6 packets."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg11m", d, packets=2e6)
        cells = read_columns(os.path.join(d, "cfg11m_cells_cellprops.dat"))
        rho = np.stack([read_columns(os.path.join(d, "cfg11m_dns_%d_rho.dat" % h))[:, 1] for h in range(2)])
        total = read_fits_cube(os.path.join(d, "cfg11m_i60_total.fits"))[0].astype(np.float64)
        out = dict(sed=read_columns(os.path.join(d, "cfg11m_i60_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg11m_i60_sedstats.dat")), component_mass_density_msun_pc3=rho,
                   cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   topology=parse_topology(os.path.join(d, "cfg11m_topo_treetop.dat")), num_packets=2e6,
                   frame_total_sum=total.sum(axis=0))
    np.savez_compressed(os.path.join(HERE, "cfg11m_ref.npz"), **out)
    print("cfg11m:", {k: np.shape(v) for k, v in out.items()})


def make_cfg14em():
    """cfg11m with explicit absorption: two dust media with different mixes, interaction points in scattering optical depth, the
    absorption optical depth interpolated along the path (MediumSystem.cpp:937-955).  Same tree and densities as cfg11m."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg14em", d, packets=2e6)
        rho = np.stack([read_columns(os.path.join(d, "cfg14em_dns_%d_rho.dat" % h))[:, 1] for h in range(2)])
        base = np.load(os.path.join(HERE, "cfg11m_ref.npz"))["component_mass_density_msun_pc3"]
        assert np.array_equal(rho, base), "cfg14em must see the inputs of cfg11m"
        total = read_fits_cube(os.path.join(d, "cfg14em_i60_total.fits"))[0].astype(np.float64)
        out = dict(sed=read_columns(os.path.join(d, "cfg14em_i60_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg14em_i60_sedstats.dat")), num_packets=2e6,
                   frame_total_sum=total.sum(axis=0))
    np.savez_compressed(os.path.join(HERE, "cfg14em_ref.npz"), **out)
    print("cfg14em:", {k: np.shape(v) for k, v in out.items()})


def make_cfg12me(packets=None, tag="cfg12me_ref"):
    """Dust emission with iterations from two dust media with different mixes (cfg4s with a second, uniform shell)."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg12me", d, packets=packets)
        cells = read_columns(os.path.join(d, "cfg12me_cells_cellprops.dat"))
        rho = np.stack([read_columns(os.path.join(d, "cfg12me_dns_%d_rho.dat" % h))[:, 1] for h in range(2)])
        rfJ = read_columns(os.path.join(d, "cfg12me_rf_J.dat"))
        prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
        sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
        dustlum = [float(x) for x in re.findall(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log)]
        conv = re.search(r"Convergence reached after (\d+) iterations", log)
        out = dict(sed=read_columns(os.path.join(d, "cfg12me_sed_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg12me_sed_sedstats.dat")),
                   J_nu_shell=shell_average(cells, rfJ[:, 1:]), absorbed_primary_lsun=np.array(prim),
                   absorbed_secondary_lsun=np.array(sec), dust_luminosity_lsun=np.array(dustlum),
                   converged_after=int(conv.group(1)) if conv else -1, num_packets=packets or 2e5)
        if packets is None:
            T = read_columns(os.path.join(d, "cfg12me_temp_dust_T.dat"))
            out.update(component_mass_density_msun_pc3=rho, cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                       topology=parse_topology(os.path.join(d, "cfg12me_topo_treetop.dat")),
                       temperature=T[:, 1].astype(np.float32))
        else:
            base = np.load(os.path.join(HERE, "cfg12me_ref.npz"))["component_mass_density_msun_pc3"]
            assert np.array_equal(rho, base), "the high-statistics run must see the inputs of the base fixture"
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag + ":", {k: np.shape(v) for k, v in out.items()}, prim, sec, conv and conv.group(0))


def make_cfg12me_hi():
    make_cfg12me(packets=2e6, tag="cfg12me_hi_ref")


def make_hi(name, packets=2e7):
    """High-statistics companion of a fixture: the same ski with `packets` histories, still with `-t 1` because the
    tree and the cell densities are sampled from the thread's random stream (Random.cpp:31-36) and must stay those of
    the base fixture; only the SED, its statistics and the wavelength-summed frame are kept."""
    with tempfile.TemporaryDirectory() as d:
        run_reference(name, d, threads=1, packets=packets)
        out = dict(sed=read_columns(os.path.join(d, name + "_i60_sed.dat")),
                   sedstats=read_columns(os.path.join(d, name + "_i60_sedstats.dat")), num_packets=packets)
        out["frame_total_sum"] = read_fits_cube(os.path.join(d, name + "_i60_total.fits"))[0].astype(np.float64).sum(axis=0)
        dens = read_columns(os.path.join(d, name + "_cells_cellprops.dat"))[:, 6]
        base = np.load(os.path.join(HERE, name + "_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(dens, base), "the high-statistics run must see the inputs of the base fixture"
    np.savez_compressed(os.path.join(HERE, name + "_hi_ref.npz"), **out)
    print(name + "_hi:", {k: np.shape(v) for k, v in out.items()})


def make_cfg1_hi():
    make_hi("cfg1")


def make_cfg2s_hi():
    make_hi("cfg2s")


def shell_average(cells, J, nshell=32, rmax=1.0):
    """Volume-weighted mean of the per-cell J_nu in `nshell` radial shells of width rmax/nshell pc (keeps the fixture small;
    the per-cell values are dominated by Monte-Carlo noise anyway)."""
    r = np.linalg.norm(cells[:, 1:4], axis=1)
    V = cells[:, 4]
    k = np.minimum((r / (rmax / nshell)).astype(int), nshell - 1)
    num = np.stack([np.bincount(k, weights=V * J[:, ell], minlength=nshell) for ell in range(J.shape[1])], axis=1)
    den = np.bincount(k, weights=V, minlength=nshell)
    return num / np.maximum(den, 1e-300)[:, None]


def make_cfg6m():
    """The multi-feature ski: every output file the reference writes, with 4e6 packets (-t 1)."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg6m", d, packets=4e6)
        out = dict(sed_sed=read_columns(os.path.join(d, "cfg6m_sed_sed.dat")),
                   sed_stats=read_columns(os.path.join(d, "cfg6m_sed_sedstats.dat")),
                   full_sed=read_columns(os.path.join(d, "cfg6m_full_sed.dat")),
                   frame_total=read_fits_cube(os.path.join(d, "cfg6m_frame_total.fits"))[0],
                   mass_density_msun_pc3=read_columns(os.path.join(d, "cfg6m_cells_cellprops.dat"))[:, 6],
                   J_nu=read_columns(os.path.join(d, "cfg6m_rf_J.dat"))[:, 1:].astype(np.float32), num_packets=4e6)
        for c in ("total", "transparent", "primarydirect", "primaryscattered"):
            out["full_" + c] = read_fits_cube(os.path.join(d, f"cfg6m_full_{c}.fits"))[0]
    np.savez_compressed(os.path.join(HERE, "cfg6m_ref.npz"), **out)
    print("cfg6m:", {k: np.shape(v) for k, v in out.items()})


def make_cfg4s():
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg4s", d)
        sed = read_columns(os.path.join(d, "cfg4s_sed_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg4s_sed_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg4s_cells_cellprops.dat"))
        topo = parse_topology(os.path.join(d, "cfg4s_topo_treetop.dat"))
        rfJ = read_columns(os.path.join(d, "cfg4s_rf_J.dat"))
        T = read_columns(os.path.join(d, "cfg4s_temp_dust_T.dat"))
        # convergence log lines, MonteCarloSimulation.cpp:193-214
        prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
        sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
        dustlum = [float(x) for x in re.findall(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log)]
        conv = re.search(r"Convergence reached after (\d+) iterations", log)
        out = dict(sed=sed, sedstats=stats, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   topology=topo, J_nu_shell=shell_average(cells, rfJ[:, 1:]), temperature=T[:, 1].astype(np.float32), absorbed_primary_lsun=np.array(prim),
                   absorbed_secondary_lsun=np.array(sec), dust_luminosity_lsun=np.array(dustlum), converged_after=int(conv.group(1)) if conv else -1,
                   num_packets=2e5)
    np.savez_compressed(os.path.join(HERE, "cfg4s_ref.npz"), **out)
    print("cfg4s:", {k: np.shape(v) for k, v in out.items()}, prim, sec, conv and conv.group(0))


def make_cfg4s_hi(packets=2e6):
    """High-statistics companion of cfg4s: the same ski (same tree and densities: -t 1) with ten times the packets per
    segment, so that the tests can compare the noisy quantities (SED bins, shell-averaged radiation field, the iteration
    log) against values whose own Monte-Carlo error is three times smaller."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg4s", d, packets=packets)
        cells = read_columns(os.path.join(d, "cfg4s_cells_cellprops.dat"))
        base = np.load(os.path.join(HERE, "cfg4s_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(cells[:, 6], base), "the high-statistics run must see the inputs of the base fixture"
        rfJ = read_columns(os.path.join(d, "cfg4s_rf_J.dat"))
        prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
        sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
        dustlum = [float(x) for x in re.findall(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log)]
        conv = re.search(r"Convergence reached after (\d+) iterations", log)
        out = dict(sed=read_columns(os.path.join(d, "cfg4s_sed_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg4s_sed_sedstats.dat")),
                   J_nu_shell=shell_average(cells, rfJ[:, 1:]), absorbed_primary_lsun=np.array(prim),
                   absorbed_secondary_lsun=np.array(sec), dust_luminosity_lsun=np.array(dustlum),
                   converged_after=int(conv.group(1)) if conv else -1, num_packets=packets)
    np.savez_compressed(os.path.join(HERE, "cfg4s_hi_ref.npz"), **out)
    print("cfg4s_hi:", {k: np.shape(v) for k, v in out.items()}, prim, sec, conv and conv.group(0))


def make_cfg7v():
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg7v", d)
        sed = read_columns(os.path.join(d, "cfg7v_sed_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg7v_sed_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg7v_cells_cellprops.dat"))
        T = read_columns(os.path.join(d, "cfg7v_temp_dust_T.dat"))
        prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
        sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
        conv = re.search(r"Convergence reached after (\d+) iterations", log)
        out = dict(sed=sed, sedstats=stats, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   temperature=T[:, 1].astype(np.float32), absorbed_primary_lsun=np.array(prim),
                   absorbed_secondary_lsun=np.array(sec), converged_after=int(conv.group(1)) if conv else -1, num_packets=2e5)
    np.savez_compressed(os.path.join(HERE, "cfg7v_ref.npz"), **out)
    print("cfg7v:", {k: np.shape(v) for k, v in out.items()}, prim, sec, conv and conv.group(0))


def make_cfg7v_hi(packets=2e6):
    """High-statistics companion of cfg7v (same sites and densities: -t 1, same seed): ten times the packets per segment."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg7v", d, packets=packets)
        cells = read_columns(os.path.join(d, "cfg7v_cells_cellprops.dat"))
        base = np.load(os.path.join(HERE, "cfg7v_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(cells[:, 6], base), "the high-statistics run must see the inputs of the base fixture"
        prim = [float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]
        sec = [float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]
        conv = re.search(r"Convergence reached after (\d+) iterations", log)
        out = dict(sed=read_columns(os.path.join(d, "cfg7v_sed_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg7v_sed_sedstats.dat")), absorbed_primary_lsun=np.array(prim),
                   absorbed_secondary_lsun=np.array(sec), converged_after=int(conv.group(1)) if conv else -1,
                   num_packets=packets)
    np.savez_compressed(os.path.join(HERE, "cfg7v_hi_ref.npz"), **out)
    print("cfg7v_hi:", {k: np.shape(v) for k, v in out.items()}, prim, sec, conv and conv.group(0))


def sph_particles(n=6000, seed=12345):
    """SURVEY.md 8d cfg5 recipe scaled down: columns x y z h M (pc, pc, pc, pc, Msun)."""
    rng = np.random.default_rng(seed)
    R = rng.gamma(2.0, 3000.0, size=4 * n)
    R = R[R < 15000.0][:n]
    phi = rng.uniform(0, 2 * np.pi, size=n)
    z = np.clip(rng.laplace(0.0, 250.0, size=n), -1900.0, 1900.0)
    h = 400.0 * (1 + R / 8000.0)
    M = np.full(n, 1e3 * 5e5 / n)   # the same total dust mass as the 5e5-particle configuration
    return np.stack([R * np.cos(phi), R * np.sin(phi), z, h, M], axis=1)


def sph_text(p):
    head = "# column 1: x (pc)\n# column 2: y (pc)\n# column 3: z (pc)\n# column 4: h (pc)\n# column 5: M (Msun)\n"
    return head + "\n".join(" ".join("%.8e" % v for v in row) for row in p) + "\n"


def make_cfg5s():
    p = sph_particles()
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg5s", d, {"sph.txt": sph_text(p)})
        sed = read_columns(os.path.join(d, "cfg5s_i60_sed.dat"))
        stats = read_columns(os.path.join(d, "cfg5s_i60_sedstats.dat"))
        cells = read_columns(os.path.join(d, "cfg5s_cells_cellprops.dat"))
        out = dict(sed=sed, sedstats=stats, particles=p, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4],
                   cell_volume_pc3=cells[:, 4], num_packets=2e5,
                   seconds=float(re.search(r"Finished primary emission in ([0-9.]+) s", log).group(1)))
        out.update(frames(d, "cfg5s_i60", ["total", "transparent", "primarydirect", "primaryscattered", "stats0",
                                           "stats1", "stats2"]))
    np.savez_compressed(os.path.join(HERE, "cfg5s_ref.npz"), **out)
    print("cfg5s:", {k: np.shape(v) for k, v in out.items()})


def make_cfg13p():
    """ParticleMedium sampled at the centres of Cartesian cells (numDensitySamples = 1): the reference's smoothed-particle density,
    deterministic."""
    p = sph_particles()
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg13p", d, {"sph.txt": sph_text(p)})
        cells = read_columns(os.path.join(d, "cfg13p_cells_cellprops.dat"))
        out = dict(particles=p, mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4])
    np.savez_compressed(os.path.join(HERE, "cfg13p_ref.npz"), **out)
    print("cfg13p:", {k: np.shape(v) for k, v in out.items()})


def make_cfg15k():
    """Kinematics: a moving point source with a narrow emission feature in an expanding dust shell, dust emission without
    iterations (tests/golden/ski/cfg15k.ski).  Three SED instruments (two opposite lines of sight along the source's velocity
    with a fine wavelength grid around the feature, one on the default grid), the radiation field on a grid that is fine around
    the feature (volume-weighted in 16 radial shells), the dust luminosity from the log."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg15k", d)
        cells = read_columns(os.path.join(d, "cfg15k_cells_cellprops.dat"))
        rfJ = read_columns(os.path.join(d, "cfg15k_rf_J.dat"))
        out = dict(mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   J_nu_shell=shell_average(cells, rfJ[:, 1:], nshell=16),
                   rf_wavelengths_micron=read_columns(os.path.join(d, "cfg15k_rf_wavelengths.dat"))[:, 0],
                   # characteristic wavelength, effective width, left and right border of every bin of the ListWavelengthGrid
                   rf_grid_micron=read_columns(os.path.join(d, "cfg15k_rf_wavelengths.dat")),
                   dust_luminosity_lsun=float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1)),
                   num_packets=1e6)
        for name in ("fwd", "bwd", "sed"):
            out["sed_" + name] = read_columns(os.path.join(d, "cfg15k_%s_sed.dat" % name))
            out["sedstats_" + name] = read_columns(os.path.join(d, "cfg15k_%s_sedstats.dat" % name))
    np.savez_compressed(os.path.join(HERE, "cfg15k_ref.npz"), **out)
    print("cfg15k:", {k: np.shape(v) for k, v in out.items()}, out["dust_luminosity_lsun"])


def make_cfg18ke():
    """cfg15k with a second, rotating dust component of another mix and explicit absorption (tests/golden/ski/cfg18ke.ski); the
    mass density of each component from the reference's DensityProbe.  4e5 packets (a history is ten scatterings long)."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg18ke", d, packets=4e5)
        cells = read_columns(os.path.join(d, "cfg18ke_cells_cellprops.dat"))
        rho = np.stack([read_columns(os.path.join(d, "cfg18ke_dns_%d_rho.dat" % h))[:, 1] for h in range(2)])
        rfJ = read_columns(os.path.join(d, "cfg18ke_rf_J.dat"))
        out = dict(component_mass_density_msun_pc3=rho, cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   J_nu_shell=shell_average(cells, rfJ[:, 1:], nshell=16),
                   rf_wavelengths_micron=read_columns(os.path.join(d, "cfg18ke_rf_wavelengths.dat"))[:, 0],
                   dust_luminosity_lsun=float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1)),
                   num_packets=4e5)
        for name in ("fwd", "bwd", "sed"):
            out["sed_" + name] = read_columns(os.path.join(d, "cfg18ke_%s_sed.dat" % name))
            out["sedstats_" + name] = read_columns(os.path.join(d, "cfg18ke_%s_sedstats.dat" % name))
    np.savez_compressed(os.path.join(HERE, "cfg18ke_ref.npz"), **out)
    print("cfg18ke:", {k: np.shape(v) for k, v in out.items()}, out["dust_luminosity_lsun"])


def make_cfg18ke_velocity():
    """The per-cell bulk velocities the reference's set-up stores for cfg18ke (VelocityProbe, per-cell form): two moving components,
    aggregated by number density (MediumSystem.cpp:330-345).  Pins the host mirror's vector fields and aggregation."""
    with tempfile.TemporaryDirectory() as d:
        text = open(os.path.join(HERE, "ski", "cfg18ke.ski")).read()
        text = text.replace('<SpatialCellPropertiesProbe probeName="cells" wavelength="0.55 micron"/>',
                            '<SpatialCellPropertiesProbe probeName="cells" wavelength="0.55 micron"/>'
                            '<VelocityProbe probeName="vel"><form type="Form"><PerCellForm/></form></VelocityProbe>')
        text = re.sub(r'numPackets="[^"]*"', 'numPackets="100"', text, count=1)
        ski = os.path.join(d, "v.ski")
        open(ski, "w").write(text)
        subprocess.check_call([SKIRT, "-t", "1", "-b", "-o", d, ski], stdout=subprocess.DEVNULL)
        v = read_columns(os.path.join(d, "v_vel_v.dat"))[:, 1:]
    np.savez_compressed(os.path.join(HERE, "cfg18ke_velocity_ref.npz"), velocity_km_s=v)
    print("cfg18ke_velocity:", v.shape, np.abs(v).max())


def make_cfg19ks():
    """A rotating ring source (GeometricSource with a CylindricalVectorField velocity) in dust at rest: tests/golden/ski/cfg19ks.ski."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg19ks", d)
        cells = read_columns(os.path.join(d, "cfg19ks_cells_cellprops.dat"))
        rfJ = read_columns(os.path.join(d, "cfg19ks_rf_J.dat"))
        out = dict(mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   J_nu_shell=shell_average(cells, rfJ[:, 1:], nshell=16),
                   rf_wavelengths_micron=read_columns(os.path.join(d, "cfg19ks_rf_wavelengths.dat"))[:, 0],
                   dust_luminosity_lsun=float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1)),
                   num_packets=1e6)
        for name in ("fwd", "bwd", "sed"):
            out["sed_" + name] = read_columns(os.path.join(d, "cfg19ks_%s_sed.dat" % name))
            out["sedstats_" + name] = read_columns(os.path.join(d, "cfg19ks_%s_sedstats.dat" % name))
    np.savez_compressed(os.path.join(HERE, "cfg19ks_ref.npz"), **out)
    print("cfg19ks:", {k: np.shape(v) for k, v in out.items()}, out["dust_luminosity_lsun"])


def make_cfg20kn():
    """cfg2s with a rotating disk (source and dust ring at 9000 km/s), a narrow emission feature and NON-forced scattering: same
    seed and set-up as cfg2s, hence its tree and densities.  2e5 packets (the reference spends 1e-3 s per history on this path)."""
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg20kn", d, packets=2e5)
        cells = read_columns(os.path.join(d, "cfg20kn_cells_cellprops.dat"))
        base = np.load(os.path.join(HERE, "cfg2s_ref.npz"))["mass_density_msun_pc3"]
        assert np.array_equal(cells[:, 6], base), "cfg20kn must see the tree and the densities of cfg2s"
        out = dict(num_packets=2e5)
        for name in ("edge", "i60"):
            out["sed_" + name] = read_columns(os.path.join(d, "cfg20kn_%s_sed.dat" % name))
            out["sedstats_" + name] = read_columns(os.path.join(d, "cfg20kn_%s_sedstats.dat" % name))
    np.savez_compressed(os.path.join(HERE, "cfg20kn_ref.npz"), **out)
    print("cfg20kn:", {k: np.shape(v) for k, v in out.items()})


def make_cfg15k_opticalprops():
    """Cross sections and asymmetry parameter of cfg15k's MeanListDustMix at 400 wavelengths as the reference's
    OpticalMaterialPropertiesProbe writes them (the probe's wavelengths join the grid of the dust property tables, so the values
    are the mix's own interpolation at exactly those wavelengths).  Pins the host mirror's MeanListDustMix."""
    with tempfile.TemporaryDirectory() as d:
        text = open(os.path.join(HERE, "ski", "cfg15k.ski")).read()
        text = text.replace('<SpatialCellPropertiesProbe probeName="cells" wavelength="0.55 micron"/>',
                            '<SpatialCellPropertiesProbe probeName="cells" wavelength="0.55 micron"/>'
                            '<OpticalMaterialPropertiesProbe probeName="opt"><wavelengthGrid type="WavelengthGrid">'
                            '<LogWavelengthGrid minWavelength="0.16 micron" maxWavelength="900 micron" numWavelengths="400"/>'
                            '</wavelengthGrid></OpticalMaterialPropertiesProbe>')
        text = re.sub(r'numPackets="[^"]*"', 'numPackets="100"', text, count=1)
        ski = os.path.join(d, "o.ski")
        open(ski, "w").write(text)
        subprocess.check_call([SKIRT, "-t", "1", "-b", "-o", d, ski], stdout=subprocess.DEVNULL)
        t = read_columns(os.path.join(d, "o_opt_opticalprops_0.dat"))
    np.savez_compressed(os.path.join(HERE, "cfg15k_opticalprops_ref.npz"), table=t)
    print("cfg15k_opticalprops:", t.shape)


def make_cfg16d():
    """Dynamic medium state: a ClearDensityRecipe carves a cavity around the source in primary emission iterations, merged primary
    and secondary iterations follow, then the regular segments (tests/golden/ski/cfg16d.ski).  The fixture holds the initial
    densities (a run of the same ski without the recipe's effect: threshold 1e30), the final ones, the per-iteration log values
    and the SED."""
    with tempfile.TemporaryDirectory() as d:
        text = open(os.path.join(HERE, "ski", "cfg16d.ski")).read()
        open(os.path.join(d, "cfg16d0.ski"), "w").write(
            text.replace('fieldStrengthThreshold="300"', 'fieldStrengthThreshold="1e30"').replace('numPackets="4e5"', 'numPackets="100"'))
        subprocess.check_call([SKIRT, "-t", "1", "-b", "-o", d, os.path.join(d, "cfg16d0.ski")], stdout=subprocess.DEVNULL)
        initial = read_columns(os.path.join(d, "cfg16d0_cells_cellprops.dat"))
        log = run_reference("cfg16d", d)
        cells = read_columns(os.path.join(d, "cfg16d_cells_cellprops.dat"))
        assert np.array_equal(initial[:, 1:5], cells[:, 1:5])
        out = dict(initial_mass_density_msun_pc3=initial[:, 6], final_mass_density_msun_pc3=cells[:, 6],
                   cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   updated_cells=np.array([int(x) for x in re.findall(r"Updated cells: (\d+) out of", log)]),
                   dust_luminosity_lsun=np.array([float(x) for x in re.findall(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log)]),
                   absorbed_primary_lsun=np.array([float(x) for x in re.findall(r"dust-absorbed primary luminosity is ([0-9.eE+-]+) Lsun", log)]),
                   absorbed_secondary_lsun=np.array([float(x) for x in re.findall(r"dust-absorbed secondary luminosity in iteration \d+ is ([0-9.eE+-]+) Lsun", log)]),
                   primary_iterations=len(re.findall(r"Finished primary emission iteration", log)),
                   merged_iterations=len(re.findall(r"Finished merged primary and secondary emission iteration", log)),
                   sed=read_columns(os.path.join(d, "cfg16d_sed_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg16d_sed_sedstats.dat")), num_packets=4e5)
    np.savez_compressed(os.path.join(HERE, "cfg16d_ref.npz"), **out)
    print("cfg16d:", {k: (np.shape(v) if np.ndim(v) > 1 else v) for k, v in out.items() if "density" not in k and "cell" not in k and "sed" not in k})


def make_cfg17c():
    """Dust heated by the CMB at redshift 6 (tests/golden/ski/cfg17c.ski): SED of the observer-frame instrument, the dust
    temperatures of the TemperatureProbe, the dust luminosity; the sampled densities as input."""
    with tempfile.TemporaryDirectory() as d:
        log = run_reference("cfg17c", d)
        cells = read_columns(os.path.join(d, "cfg17c_cells_cellprops.dat"))
        head = open(os.path.join(d, "cfg17c_sed_sed.dat")).readline()
        out = dict(mass_density_msun_pc3=cells[:, 6], cell_center_pc=cells[:, 1:4], cell_volume_pc3=cells[:, 4],
                   sed=read_columns(os.path.join(d, "cfg17c_sed_sed.dat")),
                   sedstats=read_columns(os.path.join(d, "cfg17c_sed_sedstats.dat")),
                   temperature=read_columns(os.path.join(d, "cfg17c_temp_dust_T.dat"))[:, 1].astype(np.float32),
                   dust_luminosity_lsun=float(re.search(r"Dust luminosity: ([0-9.eE+-]+) Lsun", log).group(1)),
                   redshift=6.0, luminosity_distance_mpc=float(re.search(r"luminosity distance ([0-9.eE+-]+) Mpc", head).group(1)),
                   num_packets=4e5)
    np.savez_compressed(os.path.join(HERE, "cfg17c_ref.npz"), **out)
    print("cfg17c:", {k: np.shape(v) for k, v in out.items()}, out["dust_luminosity_lsun"], out["luminosity_distance_mpc"])


def make_cfg1_formats():
    """The text headers and FITS cards of the files the reference writes for cfg1 (formats only: 1e4 packets), for the test of
    skirt9_b200/output.py."""
    import json
    with tempfile.TemporaryDirectory() as d:
        run_reference("cfg1", d, packets=1e4)
        def comments(name, limit=None):
            lines = [ln.rstrip("\n") for ln in open(os.path.join(d, name)) if ln.startswith("#")]
            return lines[:limit] if limit else lines
        def cards(name):
            raw = open(os.path.join(d, name), "rb").read()
            out, blocks, pos = [], 0, 0
            while blocks < 2:       # primary header and the header of the table extension
                c = raw[pos:pos + 80].decode("ascii")
                pos += 80
                if c.startswith("END"):
                    blocks += 1
                    out.append("END")
                    pos = (pos + 2879) // 2880 * 2880
                    if blocks == 1:
                        pos += (64 * 64 * 4 + 2879) // 2880 * 2880
                elif not c.startswith("DATE"):
                    out.append(c.rstrip())
            return out, raw[pos:pos + 16].decode("ascii")
        c, row = cards("cfg1_i60_total.fits")
        out = dict(sed=comments("cfg1_i60_sed.dat"), sedstats=comments("cfg1_i60_sedstats.dat"), rf=comments("cfg1_rf_J.dat"),
                   fits_cards=c, fits_table_row=row, files=sorted(f for f in os.listdir(d) if f.startswith("cfg1_i60_") and "stats" not in f))
    json.dump(out, open(os.path.join(HERE, "cfg1_formats.json"), "w"), indent=1)
    print("cfg1_formats:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.exists(SKIRT):
        raise SystemExit("oracle/_ref is not built: run `make -C oracle -f ref.mk -j8` where /root/reference exists")
    todo = sys.argv[1:] or ["cfg1", "cfg2s"]
    for name in todo:
        globals()["make_" + name]()
