"""Pins the oracle (CPU, `-m "not gpu"`) and the CUDA engine (`-m gpu`) against outputs of the UNMODIFIED reference.

The fixtures under tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs oracle/_ref (SKIRT 9
built from /root/reference) with `-t 1` on tests/golden/ski/*.ski and stores the reference's own per-cell densities /
tree topology (its inputs) next to its SED, frames, statistics and radiation field (its outputs).  Both sides therefore
see identical grids and densities, and differ only by the random streams (MT19937-64 in the reference, Philox here), so
agreement is statistical: |F - F_ref| <= 4 sqrt(R^2 + R_ref^2) F with R from the Sum w^k statistics the reference defines
(FluxRecorder.hpp:50-63), plus the exact (noise-free) identities.  This is SURVEY.md 8d's parity criterion.
"""
import math
import os

import numpy as np
import pytest

from skirt9_b200 import abi, configs
from skirt9_b200 import host as H
from tests.oracle_lib import OracleEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RHO = H.MSUN / H.PC ** 3  # Msun/pc3 -> kg/m3


def load(name):
    return np.load(os.path.join(GOLD, name + "_ref.npz"))


def rel_error(stats_row):
    """R = sqrt(Sum w^2/(Sum w)^2 - 1/N), FluxRecorder.hpp:50-63; stats_row = (N, Sum w, Sum w^2, ...)."""
    n, w1, w2 = stats_row[0], stats_row[1], stats_row[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.maximum(w2 / (w1 * w1) - 1.0 / np.maximum(n, 1), 0.0))


def cfg1_from_reference(num_packets):
    g = load("cfg1")
    sim = configs.cfg1(num_packets=num_packets, seed=0, record_statistics=True)
    sim.density = g["mass_density_msun_pc3"] * RHO / sim.medium.mix.MU
    sim.setup()
    return sim, g


def cfg2s_from_reference(num_packets):
    g = load("cfg2s")
    sim = configs.cfg2(num_packets=num_packets, seed=0, max_level=6, max_dust_fraction=1e-4, num_pixels=64,
                       num_wavelengths=10, record_statistics=True)
    pc = H.PC
    sim.grid = H.FileTreeSpatialGrid(-20000 * pc, 20000 * pc, -20000 * pc, 20000 * pc, -2000 * pc, 2000 * pc,
                                     g["topology"], policyOrder=True)
    sim.density = g["mass_density_msun_pc3"] * RHO / sim.medium.mix.MU
    sim.setup()
    # the renumbered tree must be the reference's: same cell centres in the same order
    boxes = sim.grid.cell_boxes()
    np.testing.assert_allclose(0.5 * (boxes[:, :3] + boxes[:, 3:]) / pc, g["cell_center_pc"], rtol=1e-8, atol=1e-6)
    np.testing.assert_allclose(sim.volume / pc ** 3, g["cell_volume_pc3"], rtol=1e-8)
    return sim, g


def check_cfg1(sim, e, g, n):
    sed = g["sed"][0]  # lambda, total, transparent, direct, scattered, ...
    tr = sim.sed_flux_density(e, 0, abi.SK_COMP_TRANSPARENT)[0]
    di = sim.sed_flux_density(e, 0, abi.SK_COMP_PRIMARY_DIRECT)[0]
    sc = sim.sed_flux_density(e, 0, abi.SK_COMP_PRIMARY_SCATTERED)[0]
    tot = sim.sed_flux_density(e, 0, abi.SK_COMP_TOTAL)[0]
    # noise-free identities: every packet has the same wavelength, position and optical depth to the observer
    assert tr == pytest.approx(sed[2], rel=1e-8)
    assert di == pytest.approx(sed[3], rel=1e-7)   # densities travel through a %.9e text file
    # Monte-Carlo part
    r_ref = rel_error(g["sedstats"][0, 1:])
    st = e.read_sed_stats(0)[:, 0]
    r_own = rel_error(st)
    assert st[0] == n
    tol = 4.0 * math.hypot(r_ref, r_own)
    assert abs(tot - sed[1]) <= tol * sed[1], (tot, sed[1], tol)
    # the scattered flux carries all the noise of the total
    assert abs(sc - sed[4]) <= tol * sed[1], (sc, sed[4], tol)
    # frames: noise-free components per pixel, the scattered frame in 8x8 blocks with the reference's own per-pixel statistics
    f_tr = sim.surface_brightness(e, 0, abi.SK_COMP_TRANSPARENT)
    f_di = sim.surface_brightness(e, 0, abi.SK_COMP_PRIMARY_DIRECT)
    f_sc = sim.surface_brightness(e, 0, abi.SK_COMP_PRIMARY_SCATTERED)
    np.testing.assert_allclose(f_tr, g["frame_transparent"], rtol=2e-6, atol=0)  # float32 files
    np.testing.assert_allclose(f_di, g["frame_primarydirect"], rtol=2e-6, atol=0)
    blk = lambda a: a.reshape(8, 8, 8, 8).sum(axis=(1, 3))
    a, b = blk(f_sc[0]), blk(g["frame_primaryscattered"][0].astype(float))
    s0, s1, s2 = (blk(g["frame_stats%d" % k][0].astype(float)) for k in range(3))
    # relative error of a block of the reference frame from its per-pixel Sum w, Sum w^2 (upper bound: pixels of one
    # history are positively correlated only through multiple scattering)
    r_blk = np.sqrt(np.maximum(s2, 1e-300)) / np.maximum(s1, 1e-300)
    scale = math.sqrt(g["num_packets"] / n)
    ok = b > 0.02 * b.max()
    assert np.all(np.abs(a - b)[ok] <= 5.0 * r_blk[ok] * math.hypot(1.0, scale) * b[ok] + 1e-3 * b.max())
    # radiation field: J per cell in radial shells (noise per cell ~ several % at 1e6 packets; per shell < 1 %)
    J = sim.mean_intensity_nu(e, 0)[:, 0]
    Jref = g["J_nu"][:, 0]
    r = np.linalg.norm(g["cell_center_pc"].astype(float), axis=1)
    shell = np.minimum((r / 0.125).astype(int), 13)
    a = np.bincount(shell, weights=J, minlength=14)
    b = np.bincount(shell, weights=Jref, minlength=14)
    np.testing.assert_allclose(a[:13], b[:13], rtol=0.02 * max(1.0, scale / 2))  # shell 13 = the 8 corner cells
    assert J.sum() == pytest.approx(Jref.sum(), rel=0.004 * max(1.0, scale))


def check_cfg2s(sim, e, g, n, nsigma=4.0):
    sed = g["sed"]
    scale = math.sqrt(g["num_packets"] / n)
    r_ref = rel_error(g["sedstats"][:, 1:].T)
    r_own = rel_error(e.read_sed_stats(0))
    lam = sim.defaultWavelengthGrid.lambdav
    np.testing.assert_allclose(lam * 1e6, sed[:, 0], rtol=1e-9)
    tol = nsigma * np.hypot(r_ref, r_own)
    for col, comp in ((1, abi.SK_COMP_TOTAL), (2, abi.SK_COMP_TRANSPARENT), (3, abi.SK_COMP_PRIMARY_DIRECT),
                      (4, abi.SK_COMP_PRIMARY_SCATTERED)):
        f = sim.sed_flux_density(e, 0, comp)
        # the statistics are those of the total flux; components are bounded by the same absolute error
        assert np.all(np.abs(f - sed[:, col]) <= tol * sed[:, 1] + 1e-12 * sed[:, 1].max()), (comp, f, sed[:, col], tol)
    # the attenuation direct/transparent per bin (launch positions are sampled, so this carries the same noise)
    di, tr = sim.sed_flux_density(e, 0, abi.SK_COMP_PRIMARY_DIRECT), sim.sed_flux_density(e, 0, abi.SK_COMP_TRANSPARENT)
    assert np.all(np.abs(di / tr - sed[:, 3] / sed[:, 2]) <= tol * sed[:, 3] / sed[:, 2])
    # frames, 8x8 blocks of the wavelength-summed total frame
    a = sim.surface_brightness(e, 0, abi.SK_COMP_TOTAL).sum(axis=0)
    b = g["frame_total"].astype(float).sum(axis=0)
    blk = lambda x: x.reshape(8, 8, 8, 8).sum(axis=(1, 3))
    a, b = blk(a), blk(b)
    ok = b > 0.05 * b.max()
    np.testing.assert_allclose(a[ok], b[ok], rtol=0.05 * max(1.0, scale))
    assert a.sum() == pytest.approx(b.sum(), rel=0.005 * max(1.0, scale))


# ---------------------------------------------------------------- CPU: the oracle against the reference
def test_oracle_matches_reference_cfg1():
    n = 100000
    sim, g = cfg1_from_reference(n)
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    check_cfg1(sim, e, g, n)
    c = e.counters()
    # SURVEY.md Appendix C (instrumented reference, 1e6 packets): 6.163 paths and 132.3 forward segments per packet
    assert c["forward_paths"] / n == pytest.approx(6.163, rel=0.01)
    assert c["forward_segments"] / n == pytest.approx(132.30, rel=0.01)
    assert c["peel_segments"] / n == pytest.approx(143.98, rel=0.01)


def test_oracle_matches_reference_cfg2s():
    n = 300000
    sim, g = cfg2s_from_reference(n)
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    # 5 sigma at this small sample: R from Sum w^k underestimates the heavy-tailed noise of the biased wavelength sampling
    check_cfg2s(sim, e, g, n, nsigma=5.0)


# ---------------------------------------------------------------- GPU: the engine against the reference
@pytest.mark.gpu
def test_engine_matches_reference_cfg1(engine_lib):
    n = 4000000
    sim, g = cfg1_from_reference(n)
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(e)
    check_cfg1(sim, e, g, n)


@pytest.mark.gpu
def test_engine_matches_reference_cfg2s(engine_lib):
    n = 4000000
    sim, g = cfg2s_from_reference(n)
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=engine_lib))
    sim.run(e)
    check_cfg2s(sim, e, g, n)
