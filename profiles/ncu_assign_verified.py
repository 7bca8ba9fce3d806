"""One pass of each assign mode at the C2 shape for an ncu launch list (gpu__time_duration per kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import sift_like
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP

dev = ops.require_cuda()
rng = np.random.default_rng(0)
m, d, k = 1_000_000, 128, int(os.environ.get("K", "4096"))
x = sift_like(rng, m, d)
xd = torch.from_numpy(x).to(dev)
c = x[rng.choice(m, k, replace=False)].copy()
c /= np.linalg.norm(c, axis=1, keepdims=True)
b = ops.prepare_operand(torch.from_numpy(c).to(dev))
a = ops.prepare_operand(xd, rows=True)
for _ in range(2):
    ops.assign_fused(xd, b, METRIC_IP, verified=False)
    ops.assign_fused(xd, b, METRIC_IP, verified=True)
    ops.assign_verified(a, b, METRIC_IP)
    ops.gemm_select(a, b, METRIC_IP, 1)
torch.cuda.synchronize()
