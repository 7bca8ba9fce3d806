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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_c_abi_collectives_single_process_two_gpus():
    """ise_comm_init_all / ise_allreduce_sum_f32 / ise_allgather: the exchange steps of the sharded path for a host
    that binds libise directly and drives all GPUs from ONE process (run in a subprocess: NCCL communicators of a
    single-process group and torch's own do not share a process here)."""
    code = r"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from image_search_engine_b200 import _lib
lib = _lib.load()
comm = C.c_void_p()
_lib.check(lib.ise_comm_init_all(2, None, C.byref(comm)))
assert lib.ise_comm_size(comm) == 2
k, d = 1000, 128
bufs = [torch.full((k * d + k,), float(i + 1), device=f"cuda:{i}") for i in range(2)]
ptrs = (C.c_void_p * 2)(*[b.data_ptr() for b in bufs])
streams = (C.c_void_p * 2)(*[torch.cuda.current_stream(i).cuda_stream for i in range(2)])
_lib.check(lib.ise_allreduce_sum_f32(comm, ptrs, bufs[0].numel(), streams))
for i in range(2):
    torch.cuda.synchronize(i)
    assert bool((bufs[i] == 3.0).all())
send = [torch.arange(24, dtype=torch.uint8, device=f"cuda:{i}") + 100 * i for i in range(2)]
recv = [torch.zeros(48, dtype=torch.uint8, device=f"cuda:{i}") for i in range(2)]
sp = (C.c_void_p * 2)(*[t.data_ptr() for t in send]); rp = (C.c_void_p * 2)(*[t.data_ptr() for t in recv])
_lib.check(lib.ise_allgather(comm, sp, rp, 24, streams))
want = torch.cat([send[0].cpu(), send[1].cpu()])
for i in range(2):
    torch.cuda.synchronize(i)
    assert torch.equal(recv[i].cpu(), want)
lib.ise_comm_destroy(comm)
print("C_ABI_COLLECTIVES_OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       cwd=str(Path(__file__).resolve().parent.parent))
    assert r.returncode == 0 and "C_ABI_COLLECTIVES_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
