"""Minimal driver for ncu: the single-pass row preparation and the sorted-gather k-means update at C2 size."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import sift_like, C2
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
dev = ops.require_cuda()
rng = np.random.default_rng(2)
n, d, k = C2["n_desc"], C2["d"], C2["k"]
X = torch.from_numpy(sift_like(rng, n, d)).to(dev)
cent = X[torch.randperm(n, device=dev)[:k]].clone()
ops.normalize_l2_(cent)
a = ops.compact_operand(ops.prepare_operand(X, rows=True))
b = ops.prepare_operand(cent)
_, words = ops.gemm_select(a, b, METRIC_IP, 1)
accum = torch.zeros((k * d + k,), dtype=torch.float32, device=dev)
sums, counts = accum[: k * d].view(k, d), accum[k * d:]
obj = torch.zeros((1,), dtype=torch.float64, device=dev)
ws = None
for _ in range(2):
    ops.prepare_operand(X, rows=True)
    ws = ops.kmeans_accumulate_sorted(X, words, sums, counts, obj, centroids=cent, workspace=ws)
    ws = ops.kmeans_accumulate_sorted(X, words, sums, counts, obj, centroids=cent, workspace=ws, exact_op=a)
torch.cuda.synchronize()
