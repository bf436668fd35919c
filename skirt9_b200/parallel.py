"""Multi-GPU execution of the life cycle: static sharding of history indices plus the two reductions of the reference.

The reference distributes histories over MPI ranks through a chunk server (MultiHybridParallel.cpp:26-144) and keeps a
full replica of grid, medium and tallies on every rank; after each segment it sums the radiation field over the ranks
(MediumSystem::communicateRadiationField -> ProcessManager::sumToAll, MediumSystem.cpp:1304-1313) and before output it
sums the detector arrays (FluxRecorder::calibrateAndWrite -> ProcessManager::sumToRoot, FluxRecorder.cpp:487-493).
Here one process drives one GPU; the random streams are keyed by history index (Philox), so a STATIC partition into
contiguous blocks gives the same tallies as any dynamic one and no chunk server is needed; the two reductions become
in-place all-reduces on the engine's device buffers (NCCL over NVLink), enqueued on the engine's own stream.
"""
from __future__ import annotations


def history_block(num_packets: int, rank: int, world: int):
    """Histories [first, first+count) of rank `rank`: contiguous blocks, sizes differing by at most one."""
    first = num_packets * rank // world
    last = num_packets * (rank + 1) // world
    return first, last - first


class Comm:
    """The communicator a simulation is run with.  `dist` is torch.distributed (initialised: nccl for the engine, gloo
    for the CPU tests) or None for a single rank, in which case every method is a no-op."""

    def __init__(self, dist=None):
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1

    def block(self, num_packets):
        return history_block(int(num_packets), self.rank, self.world)

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
