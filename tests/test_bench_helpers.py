"""bench.py reads the reference's (and the drop-in binary's) timings from the log files they write; these are the parsers."""
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
