"""Sharded filter on 2 GPUs == single-context filter (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_two_rank_filter_matches_single_context(built):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "scripts", "sharded_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and all("SHARDED_CHECK %s OK" % k in r.stdout for k in ("peer", "nccl", "maximal", "layout")), \
        r.stdout[-3000:]
