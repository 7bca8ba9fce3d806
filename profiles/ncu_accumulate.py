"""Minimal driver for ncu: k-means accumulate at C2 size, privatised and plain kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from bench import sift_like, C2
from image_search_engine_b200 import ops
dev = ops.require_cuda()
rng = np.random.default_rng(2)
n, d, k = C2["n_desc"], C2["d"], C2["k"]
X = torch.from_numpy(sift_like(rng, n, d)).to(dev)
cent = X[torch.randperm(n, device=dev)[:k]].clone()
ops.normalize_l2_(cent)
words = torch.randint(0, k, (n,), device=dev)
accum = torch.zeros((k * d + k,), dtype=torch.float32, device=dev)
sums, counts = accum[: k * d].view(k, d), accum[k * d:]
obj = torch.zeros((1,), dtype=torch.float64, device=dev)
ws = None
for dt_name, Xin in (("f32", X), ("u8", X.to(torch.uint8))):
    for _ in range(3):
        ws = ops.kmeans_accumulate_sorted(Xin, words, sums, counts, obj, centroids=cent, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ws = ops.kmeans_accumulate_sorted(Xin, words, sums, counts, obj, centroids=cent, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print("sorted", dt_name, "%.1f us" % (e0.elapsed_time(e1) / 5 * 1e3))
for mode in ("priv", "plain"):
    if mode == "plain":
        os.environ["ISE_ACCUMULATE_PLAIN"] = "1"
    for _ in range(3):
        ops.kmeans_accumulate(X, words, None, sums, counts, obj, centroids=cent)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.kmeans_accumulate(X, words, None, sums, counts, obj, centroids=cent)
    e1.record(); torch.cuda.synchronize()
    print(mode, "%.1f us" % (e0.elapsed_time(e1) / 5 * 1e3))
