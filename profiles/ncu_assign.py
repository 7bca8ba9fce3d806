"""Minimal driver for ncu: the verified fused assign (float32 rows converted inside gemm_select, one-product pass +
compact split re-run) on the C2 shape."""
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
b = ops.prepare_operand(cent)
for _ in range(4):
    got = ops.assign_fused(X, b, METRIC_IP)
assert got is not None
torch.cuda.synchronize()
print("ok")
