"""bench.py reads the reference's (and the drop-in binary's) timings from the log files they write; these are the parsers."""
import os

import pytest

import bench

LOG = """17/10/2026 23:59:58.100   Starting setup...
17/10/2026 23:59:58.379   Constructing the spatial tree grid...
17/10/2026 23:59:58.380   Subdividing level 0: 1 nodes
17/10/2026 23:59:59.429   Finished construction of the spatial tree grid
17/10/2026 23:59:59.429   Determining medium properties for 16626 cells...
17/10/2026 23:59:59.464   Done determining medium properties
17/10/2026 23:59:59.466 - Finished setup in 1.4 s.
17/10/2026 23:59:59.652   Starting primary emission...
17/10/2026 23:59:59.652   Launching 1e4 primary emission photon packets
18/10/2026 00:00:01.849 - Finished primary emission in 2.2 s.
"""


def test_emission_time_from_time_stamps_across_midnight():
    assert abs(bench.emission_seconds(LOG) - 2.197) < 1e-6


def test_emission_time_falls_back_to_the_timelogger_text():
    text = "18/10/2026 00:00:01.849 - Finished primary emission in 2.2 s.\n"
    assert bench.emission_seconds(text) == 2.2


def test_reference_setup_times():
    t = bench.reference_setup_times(LOG, 8)
    assert t["threads"] == 8 and t["cells"] == 16626
    assert abs(t["construct_tree_s"] - 1.050) < 1e-6
    assert abs(t["medium_properties_s"] - 0.035) < 1e-6
    assert bench.reference_setup_times("no such lines", 1) == {}


@pytest.mark.skipif(not os.path.exists(bench.REF_EXE), reason="oracle/_ref is built only where /root/reference exists")
def test_parity_block_adopts_the_reference_setup(tmp_path):
    """bench.py's parity block runs the engine on the set-up of the reference run it is compared with: the densities the
    reference's SpatialCellPropertiesProbe wrote become the model's, cell for cell (cfg1: Cartesian grid, 32768 cells)."""
    import subprocess
    import numpy as np
    from skirt9_b200 import host as H
    text = bench.ski_text("cfg1", 1e3, statistics=True)
    ski = tmp_path / "cfg1.ski"
    ski.write_text(text)
    subprocess.check_call([bench.REF_EXE, "-t", "2", "-b", "-o", str(tmp_path), str(ski)], stdout=subprocess.DEVNULL)
    sim = bench.make_sim("cfg1", 1e3, statistics=True)
    own = np.array(sim.density, copy=True)
    what = bench.adopt_reference_setup(sim, "cfg1", str(tmp_path))
    assert "32768 cells" in what and "sampled densities" in what
    cells = np.loadtxt(tmp_path / "cfg1_cells_cellprops.dat", comments="#")
    np.testing.assert_allclose(sim.density, cells[:, 6] * (H.MSUN / H.PC ** 3) / sim.medium.mix.MU, rtol=1e-12)
    assert not np.array_equal(sim.density, own)     # (the mirror's own sampling draws from numpy's generator)
    assert np.mean(np.abs(sim.density - own)) < 0.05 * own.mean()
