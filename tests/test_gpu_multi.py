"""Two-rank sharded update over the NVLink peer mailboxes and over NCCL (skipped on single-GPU boxes): the sharded result
must reproduce the un-sharded one (tools/multi_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("modes", ["p2p", "nccl"])
def test_two_rank_sharded_update_reproduces_single_gpu(modes):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "multi_gpu_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PARITY_CFG="small", PARITY_MODES=modes, PARITY_STEPS="10"))
    assert out.returncode == 0 and "OK" in out.stdout and "MISMATCH" not in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
