"""Minimal driver for ncu: dense and CSR histogram kernels at C2 size."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import C2
from image_search_engine_b200 import ops
dev = ops.require_cuda()
n, k, n_img = C2["n_desc"], C2["k"], C2["n_img"]
words = torch.randint(0, k, (n,), device=dev)
off = torch.arange(0, n + 1, C2["per_img"], dtype=torch.int64, device=dev)
out = torch.empty((n_img, k), dtype=torch.float64, device=dev)
for _ in range(3):
    ops.bovw_histogram(words, off, k, okapi=True, out=out)
    ops.bovw_histogram_csr(words, off, k, okapi=True)
torch.cuda.synchronize()
print("ok")
