"""Minimal driver for ncu: a few split top-1 gemm_select launches on the C2 shape."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import sift_like, C2
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
dev = ops.require_cuda()
rng = np.random.default_rng(2)
X = torch.from_numpy(sift_like(rng, C2["n_desc"], C2["d"])).to(dev)
cent = X[torch.randperm(X.shape[0], device=dev)[:C2["k"]]].clone()
ops.normalize_l2_(cent)
a = ops.compact_operand(ops.prepare_operand(X)); b = ops.prepare_operand(cent)
for _ in range(4):
    ops.gemm_select(a, b, METRIC_IP, 1)
torch.cuda.synchronize()
print("ok")
