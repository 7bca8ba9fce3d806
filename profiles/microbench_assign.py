"""Micro-benchmark of the assign (top-1) variants on the C2 shape: 1M x 128 SIFT-like rows, k = 4096."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bench import sift_like, C2
from image_search_engine_b200 import FaissKMeans, ops
from image_search_engine_b200._lib import METRIC_IP

dev = ops.require_cuda()
X = torch.from_numpy(sift_like(np.random.default_rng(2), C2["n_desc"], C2["d"])).to(dev)
km = FaissKMeans(C2["k"], n_init=1, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
km.fit(X)
cent = torch.from_numpy(km.cluster_centers_).to(dev)
a = ops.prepare_operand(X); b = ops.prepare_operand(cent)
hi = lambda o: ops.Operand(o.hi, None, o.norms, o.meta, o.n, o.d, o.ldp)
rows = torch.empty(a.n, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, fn in [
    ("prepare_operand(X)          ", lambda: ops.prepare_operand(X)),
    ("split  top-1 (hi*hi + hi*lo)", lambda: ops.gemm_select(a, b, METRIC_IP, 1)),
    ("coarse top-1                ", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 1)),
    ("coarse top-1 + flags        ", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 1, flags=(rows, cnt))),
    ("search_topk verified (ids)  ", lambda: ops.search_topk(X, a, cent, b, METRIC_IP, 1, need_distances=False)),
    ("search_topk verified (+dis) ", lambda: ops.search_topk(X, a, cent, b, METRIC_IP, 1)),
    ("search_topk split (ids)     ", lambda: ops.search_topk(X, a, cent, b, METRIC_IP, 1, precision="split", need_distances=False)),
]:
    print(f"{name}: {t(fn):7.3f} ms", flush=True)
ops.gemm_select(hi(a), hi(b), METRIC_IP, 1, flags=(rows, cnt))
print("flagged rows:", int(cnt.item()), "of", a.n, " stats:", ops.last_search_stats)
