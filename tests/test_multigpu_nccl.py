"""The sharded life cycle over NCCL on real GPUs (needs at least two): see scripts/nccl_parity.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _num_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs")
def test_two_ranks_over_nccl_equal_one_rank():
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "scripts", "nccl_parity.py")], capture_output=True, text=True, timeout=600)
    assert "NCCL PARITY PASS world 2" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
