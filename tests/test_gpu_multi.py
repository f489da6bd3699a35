"""Multi-GPU (NCCL) parity: skipped on single-GPU boxes; the host-side plumbing is covered on CPU by
tests/test_multigpu_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_pipeline_matches_single_gpu(gpu):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", os.path.join(ROOT, "scripts", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "equal=True" in r.stdout and "knn world=2 equal=True" in r.stdout, r.stdout[-2000:]
