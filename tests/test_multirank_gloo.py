"""The N>1 path on CPU: two processes over gloo run the sharded simulation driver (skirt9_b200.parallel.Comm) with the
oracle engine and must reproduce the single-rank tallies.  This covers the host logic of the multi-GPU path -- interleaved
history blocks, the all-reduce of the radiation field inside the secondary-emission iteration loop, the final all-reduce of
the detector arrays -- without a GPU; the same code drives the CUDA engine over NCCL in bench.py."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from skirt9_b200 import abi, parallel
from tests import models
from tests.oracle_lib import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from skirt9_b200 import abi, parallel
from tests import models
from tests.oracle_lib import OracleEngine

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
comm = parallel.Comm(dist, block=256)   # small blocks: the test models have a few thousand histories
sim = models.{model}.setup()
e = sim.configure(OracleEngine(sim.config_struct()))
first, count = comm.block(e, int(sim.numPackets))
sim.run(e, comm=comm)
# the oracle keeps one allocation per detector array: reduce what was read back (the CUDA engine reduces its contiguous block)
import torch
out = {{}}
for c in {comps!r}:
    t = torch.from_numpy(e.read_sed(0, c).copy())
    dist.all_reduce(t)
    out["sed%d" % c] = t.numpy()
if sim.storeRadiationField:
    out["rf1"] = e.read_rf(0)
if sim.dustEmissionWLG is not None:
    out["rf2"] = e.read_rf(1)
    out["conv"] = np.array([[c["dust_luminosity"], c["absorbed_primary"], c["absorbed_secondary"]] for c in sim.convergence])
out["block"] = np.array([first, count, e.counters()["packets"]])
if comm.rank == 0:
    np.savez({out!r}, **out)
dist.barrier()
dist.destroy_process_group()
'''


def run_two_ranks(model, comps, port):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "out.npz")
        script = os.path.join(d, "worker.py")
        open(script, "w").write(WORKER.format(root=ROOT, port=port, model=model, comps=comps, out=out))
        procs = [subprocess.Popen([sys.executable, script, str(r)], cwd=ROOT) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        return dict(np.load(out))


def test_interleaved_shares_tile_the_range():
    for n in (1, 7, 1000, 16384 * 3 + 5, 10**9 + 7):
        for world in (1, 2, 3, 8):
            for block in (256, 16384):
                shares = [parallel.interleaved_count(n, r, world, block) for r in range(world)]
                assert sum(shares) == n
                assert max(shares) - min(shares) <= block
                # the same by enumeration
                if n <= 100000:
                    idx = np.arange(n)
                    assert shares == [int(((idx // block) % world == r).sum()) for r in range(world)]


def test_history_blocks_tile_the_range():
    for n in (1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            blocks = [parallel.history_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_ranks_primary_emission_matches_single_rank():
    comps = [abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED]
    got = run_two_ranks("small_cartesian(num_packets=6001)", comps, 29611)
    sim = models.small_cartesian(num_packets=6001).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    # every rank is handed the whole segment and runs its interleaved share: rank 0 the blocks 0, 2, 4, ...
    assert list(got["block"][:2]) == [0, 6001]
    assert got["block"][2] <= parallel.interleaved_count(6001, 0, 2, 256) < 6001
    for c in comps:
        np.testing.assert_allclose(got["sed%d" % c], e.read_sed(0, c), rtol=1e-11)
    ref = e.read_rf(0)
    np.testing.assert_allclose(got["rf1"], ref, rtol=1e-10, atol=1e-12 * ref.max())


def test_two_ranks_dust_emission_iterations_match_single_rank():
    """The radiation field is all-reduced inside the iteration loop, so both ranks prepare the same secondary sources."""
    comps = [abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_SECONDARY_DIRECT, abi.SK_COMP_SECONDARY_SCATTERED]
    got = run_two_ranks("small_dust_emission(num_packets=6000)", comps, 29612)
    sim = models.small_dust_emission(num_packets=6000).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    conv = np.array([[c["dust_luminosity"], c["absorbed_primary"], c["absorbed_secondary"]] for c in sim.convergence])
    assert got["conv"].shape == conv.shape
    # summation order differs between one and two ranks, and a cumulative launch weight that lands within rounding of a
    # half-integer can move one history to the neighbouring cell: equal to a part in 1e4, not bit for bit
    np.testing.assert_allclose(got["conv"], conv, rtol=2e-4)
    for c in comps:
        a, b = got["sed%d" % c], e.read_sed(0, c)
        np.testing.assert_allclose(a, b, rtol=2e-3, atol=1e-6 * b.max())
    np.testing.assert_allclose(got["rf1"], e.read_rf(0), rtol=1e-10, atol=1e-12 * e.read_rf(0).max())


def test_two_ranks_build_the_same_grid_without_communication():
    """Engine-side set-up (octree by the density policy, sampled densities) draws from Philox streams keyed by node and
    cell index, so every rank builds the identical replica on its own: no broadcast of the grid, unlike the reference,
    whose ranks share the evaluation and sum the flags (DensityTreePolicy.cpp:283, ProcessManager::sumToAll)."""
    comps = [abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED]
    got = run_two_ranks("small_octree_engine_setup(num_packets=6000)", comps, 29613)
    sim = models.small_octree_engine_setup(num_packets=6000).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    assert list(got["block"][:2]) == [0, 6000]
    for c in comps:
        a, b = got["sed%d" % c], e.read_sed(0, c)
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-14 * b.max())


def test_two_ranks_dynamic_state_iterations_match_single_rank():
    """Primary and merged iterations over a dynamic medium state: the radiation field is all-reduced before the recipe looks at
    it, so both ranks clear the same cells and hand the same densities to their engines."""
    comps = [abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED, abi.SK_COMP_SECONDARY_DIRECT]
    got = run_two_ranks("small_dynamic_state(num_packets=8000)", comps, 29614)
    sim = models.small_dynamic_state(num_packets=8000).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    assert sum(it["updated_cells"] for it in sim.primary_iterations) > 0
    conv = np.array([[c["dust_luminosity"], c["absorbed_primary"], c["absorbed_secondary"]] for c in sim.convergence])
    assert got["conv"].shape == conv.shape
    np.testing.assert_allclose(got["conv"], conv, rtol=2e-4)
    for c in comps:
        a, b = got["sed%d" % c], e.read_sed(0, c)
        np.testing.assert_allclose(a, b, rtol=2e-3, atol=1e-6 * b.max())


def test_two_ranks_kinematics_match_single_rank():
    comps = [abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED]
    got = run_two_ranks("with_kinematics(models.two_sources_three_instruments(num_packets=6000))", comps, 29615)
    sim = models.with_kinematics(models.two_sources_three_instruments(num_packets=6000)).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    sim.run(e)
    for c in comps:
        np.testing.assert_allclose(got["sed%d" % c], e.read_sed(0, c), rtol=1e-11)
    ref = e.read_rf(0)
    np.testing.assert_allclose(got["rf1"], ref, rtol=1e-10, atol=1e-12 * ref.max())
