"""Multi-GPU execution of the life cycle: static sharding of history indices plus the two reductions of the reference.

The reference distributes histories over MPI ranks through a chunk server (MultiHybridParallel.cpp:26-144) and keeps a
full replica of grid, medium and tallies on every rank; after each segment it sums the radiation field over the ranks
(MediumSystem::communicateRadiationField -> ProcessManager::sumToAll, MediumSystem.cpp:1304-1313) and before output it
sums the detector arrays (FluxRecorder::calibrateAndWrite -> ProcessManager::sumToRoot, FluxRecorder.cpp:487-493).
Here one process drives one GPU; the random streams are keyed by history index (Philox), so a STATIC partition into
interleaved blocks of 16384 histories gives the same tallies as any dynamic one and no chunk server is needed; the two reductions become
in-place all-reduces on the engine's device buffers (NCCL over NVLink), enqueued on the engine's own stream.
"""
from __future__ import annotations


INTERLEAVE_BLOCK = 16384   # histories per block of the interleaved sharding (a power of two)


def history_block(num_packets: int, rank: int, world: int):
    """Histories [first, first+count) of rank `rank` in a CONTIGUOUS sharding: blocks whose sizes differ by at most one.
    (Kept for callers that want it; the drivers below shard by interleaved blocks, see Comm.block.)"""
    first = num_packets * rank // world
    last = num_packets * (rank + 1) // world
    return first, last - first


def interleaved_count(num_packets: int, rank: int, world: int, block: int = INTERLEAVE_BLOCK):
    """How many of the histories [0, num_packets) belong to `rank` when every world-th block of `block` is its own."""
    cycle = block * world
    rem = num_packets % cycle
    return num_packets // cycle * block + min(max(rem - rank * block, 0), block)


class Comm:
    """The communicator a simulation is run with.  `dist` is torch.distributed (initialised: nccl for the engine, gloo
    for the CPU tests) or None for a single rank, in which case every method is a no-op."""

    def __init__(self, dist=None, block=INTERLEAVE_BLOCK):
        self.interleave_block = int(block)
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1

    def block(self, engine, num_packets):
        """The range every rank hands to run_segment: the WHOLE segment, of which the engine runs its interleaved share
        (sk_engine_set_history_interleave).  The reference's chunk server hands out chunks of history indices dynamically
        (MultiHybridParallel.cpp:26-144); with counter-based random streams a static assignment gives the same tallies,
        and small interleaved blocks keep the ranks balanced where contiguous shares do not: the histories of a secondary
        emission segment are ordered by cell (DustSecondarySource.cpp:133-145)."""
        engine.set_history_interleave(self.interleave_block, self.world, self.rank)
        return 0, int(num_packets)

    def _all_reduce(self, engine, which):
        t = engine.device_tensor(which)
        if t.numel() == 0:
            return
        if t.is_cuda:
            import torch
            stream = torch.cuda.ExternalStream(engine.cuda_stream(), device=t.device)
            with torch.cuda.stream(stream):   # ordered after the segment's kernels, no host synchronisation
                self.dist.all_reduce(t)
            stream.synchronize()
        else:
            self.dist.all_reduce(t)

    def allreduce_rf(self, engine, primary):
        """ProcessManager::sumToAll on _rf1 (primary) or _rf2c (secondary), MediumSystem.cpp:1307,1310."""
        if self.dist:
            self._all_reduce(engine, 0 if primary else 2)

    def allreduce_detectors(self, engine):
        """ProcessManager::sumToRoot on every detector and statistics array, FluxRecorder.cpp:487-493 (as an all-reduce,
        so that every rank can write output)."""
        if self.dist:
            from .abi import SK_ERR_UNSUPPORTED, SkError
            try:
                self._all_reduce(engine, 3)
                self._all_reduce(engine, 4)
            except SkError as ex:
                # the test-only CPU oracle keeps one allocation per detector array and exposes no block; its tests
                # reduce the arrays after reading them back
                if ex.code != SK_ERR_UNSUPPORTED:
                    raise
