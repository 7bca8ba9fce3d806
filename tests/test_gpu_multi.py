"""Real multi-GPU parity (NCCL), skipped on boxes with one GPU: run with `gpurun --gpus 2`."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_kmeans_and_index_on_two_gpus():
    script = Path(__file__).parent / "multi_gpu_check.py"
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTI_GPU_PARITY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
