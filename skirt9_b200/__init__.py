"""skirt9_b200 -- B200-native photon-packet life-cycle engine behind the SKIRT 9 simulation-item API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/sk_engine.h),
`abi.py` (ctypes binding), `host.py` (host-side mirror of the reference's setup classes), `parallel.py`
(history sharding + the reference's two reductions over torch.distributed) and `configs.py`
(the BASELINE.json workloads).  There is no CPU implementation of the life cycle in this package.
"""
from . import abi, host, parallel  # noqa: F401

__all__ = ["abi", "host", "parallel"]
