"""N>1 path on real GPUs: launches tools/multi_gpu_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_row_sharded_solvers_match_single_gpu(exchange):
    """exchange = nccl: ncclAllReduce of the n-vector (default); p2p: the one-shot NVLink peer-memory all-reduce fused
    with the partial sum of the one-pass kernel (csrc/rls_p2p.cu, RLS_P2P=1)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the single-GPU box runs the gloo composition test on CPU instead)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, RLS_P2P="1" if exchange == "p2p" else "0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
