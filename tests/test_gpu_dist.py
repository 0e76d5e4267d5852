"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on the 1-GPU box): the row-sharded pipeline
(dist.classic_sharded under torchrun + NCCL) equals the single-GPU pipeline — D1/D2/D3_new shards
bit-exact, same sweep count, survivor CSR and walk identical."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["rows", "sym"])
@pytest.mark.parametrize("n,h,w,fs,stride", [(700, 16, 16, 40, 4), (333, 12, 20, 16, 1), (2051, 8, 8, 40, 4)])
def test_row_sharded_equals_single_gpu(n, h, w, fs, stride, mode):
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if g < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), str(h), str(w), str(fs), str(stride), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
