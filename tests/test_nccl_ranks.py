"""One process per GPU over NCCL with peer-to-peer (NVLink) halos: needs >= 2 GPUs, skipped otherwise.  The same
algorithm is covered on one GPU by tests/test_multi_rank.py (rank-threads) and on the CPU by tests/test_gloo_ranks.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_p2p_ranks_against_oracle(world):
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "dist_check.py"), "48"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for r in range(world):
        assert ("rank %d ok" % r) in p.stdout
