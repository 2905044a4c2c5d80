"""Two-rank sharded update over NCCL (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tuning", [0, 8])  # 8 = TUNE_PEER_REPLICATED: every CTA pulls the peer-reduced buffer
def test_two_rank_sharded_update_matches_oracle(tuning):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if tuning and os.environ.get("ESIKF_EXPERIMENTAL") != "1":
        pytest.skip("TUNE_PEER_REPLICATED has not run on a multi-GPU box yet (ESIKF_EXPERIMENTAL=1 enables)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(29541 + tuning),
           os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, ESIKF_TUNING=str(tuning)))
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
