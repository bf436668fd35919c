"""The output files written from the Python host mirror (skirt9_b200/output.py) have the reference's formats: text headers
and FITS cards equal to those of the files the unmodified reference writes for the same ski (tests/golden/cfg1_formats.json,
made by tests/golden/make_golden.py cfg1_formats), values equal to the calibrated arrays.  CPU only (oracle engine)."""
import json
import os

import numpy as np

from skirt9_b200 import abi, configs, output
from tests.oracle_lib import OracleEngine
from tests.skirt_files import read_columns, read_fits_cube

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def comments(path):
    return [ln.rstrip("\n") for ln in open(path) if ln.startswith("#")]


def header_cards(path):
    raw = open(path, "rb").read()
    out, blocks, pos = [], 0, 0
    while blocks < 2:
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
    return out, raw[pos:pos + 16].decode("ascii"), len(raw)


def test_cfg1_files_have_the_reference_formats(tmp_path):
    fmt = json.load(open(os.path.join(GOLD, "cfg1_formats.json")))
    sim = configs.cfg1(num_packets=20000, seed=0, record_statistics=True)
    sim.setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    prefix = str(tmp_path / "cfg1")
    paths = output.write_all(sim, e, prefix)
    names = sorted(os.path.basename(p) for p in paths)
    assert [n for n in names if n.startswith("cfg1_i60_") and "stats" not in n] == fmt["files"]
    # text files: header lines character by character, values to the 10 digits written
    assert comments(prefix + "_i60_sed.dat") == fmt["sed"]
    assert comments(prefix + "_i60_sedstats.dat") == fmt["sedstats"]
    assert comments(prefix + "_rf_J.dat") == fmt["rf"]
    sed = read_columns(prefix + "_i60_sed.dat")
    assert sed.shape == (1, 8) and sed[0, 0] == 0.55
    for col, comp in ((1, abi.SK_COMP_TOTAL), (2, abi.SK_COMP_TRANSPARENT), (3, abi.SK_COMP_PRIMARY_DIRECT), (4, abi.SK_COMP_PRIMARY_SCATTERED)):
        np.testing.assert_allclose(sed[:, col], sim.sed_flux_density(e, 0, comp), rtol=1e-9)
    assert np.all(sed[:, 5:] == 0)
    st = read_columns(prefix + "_i60_sedstats.dat")
    np.testing.assert_allclose(st[0, 1:], e.read_sed_stats(0)[:, 0], rtol=1e-9)
    J = read_columns(prefix + "_rf_J.dat")
    assert J.shape == (32 ** 3, 2) and np.array_equal(J[:, 0], np.arange(32 ** 3))
    np.testing.assert_allclose(J[:, 1], sim.mean_intensity_nu(e, 0)[:, 0], rtol=1e-9)
    # FITS: the reference's cards (all but DATE), the table extension, the pixels
    cards, row, size = header_cards(prefix + "_i60_total.fits")
    assert cards == fmt["fits_cards"]
    assert row == fmt["fits_table_row"] and size == 25920
    for name, comp in (("total", abi.SK_COMP_TOTAL), ("transparent", abi.SK_COMP_TRANSPARENT),
                       ("primarydirect", abi.SK_COMP_PRIMARY_DIRECT), ("primaryscattered", abi.SK_COMP_PRIMARY_SCATTERED)):
        cube, c = read_fits_cube("%s_i60_%s.fits" % (prefix, name))
        assert cube.shape == (1, 64, 64) and c["BUNIT"] == "MJy/sr"
        np.testing.assert_allclose(cube, sim.surface_brightness(e, 0, comp).astype(np.float32), rtol=1e-6)


def test_multi_wavelength_cube_and_its_table_extension(tmp_path):
    """A FullInstrument with several wavelength bins, scattering levels and a FrameInstrument without components: one cube per
    component, the wavelengths in the ASCII table extension, the SED columns of the scattering levels."""
    from tests import models
    sim = models.two_sources_three_instruments(num_packets=4000)
    sim.setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    prefix = str(tmp_path / "m")
    paths = [os.path.basename(p) for p in output.write_all(sim, e, prefix)]
    assert "m_sed_sed.dat" in paths and "m_sed_sedstats.dat" in paths and "m_f1_total.fits" in paths
    assert "m_f1_transparent.fits" not in paths            # (f1 does not record components)
    assert "m_f2_primaryscattered.fits" in paths and "m_rf_J.dat" in paths
    # the aperture SED instrument records two scattering levels: the wavelength, the 7 flux columns, then the two levels
    head = comments(prefix + "_sed_sed.dat")
    assert head[0] == "# SED at inclination 30 deg, azimuth 40 deg, distance 1 Mpc"
    assert head[-2:] == ["# column 9: 1-times scattered primary flux; F_nu (Jy)", "# column 10: 2-times scattered primary flux; F_nu (Jy)"]
    sed = read_columns(prefix + "_sed_sed.dat")
    g = sim.defaultWavelengthGrid
    assert sed.shape == (g.num_bins, 10)
    np.testing.assert_allclose(sed[:, 0], g.lambdav * 1e6, rtol=1e-9)
    np.testing.assert_allclose(sed[:, 8], sim.sed_flux_density(e, 0, abi.SK_COMP_PRIMARY_SCATTERED_LEVEL), rtol=1e-9)
    cube, cards = read_fits_cube(prefix + "_f2_total.fits")
    assert cube.shape == (g.num_bins, 4, 4) and int(cards["NAXIS3"]) == g.num_bins
    np.testing.assert_allclose(cube, sim.surface_brightness(e, 2, abi.SK_COMP_TOTAL).astype(np.float32), rtol=1e-6)
    assert float(cards["CROTA1"]) == 100.0 and float(cards["CROTA3"]) == 15.0
    # the table extension: one row of 16 characters per wavelength, after the padded cube
    raw = open(prefix + "_f2_total.fits", "rb").read()
    pos = 2880 + (g.num_bins * 16 * 4 + 2879) // 2880 * 2880 + 2880
    rows = [float(raw[pos + 16 * k:pos + 16 * (k + 1)]) for k in range(g.num_bins)]
    np.testing.assert_allclose(rows, g.lambdav * 1e6, rtol=1e-9)
    assert len(raw) % 2880 == 0
