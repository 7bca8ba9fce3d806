"""Is the d=128 top-1 tile bound by TMEM reads or by the event path?  Column 0 dominates every row,
so the running best never changes after the first chunk: the epilogue is a pure TMEM scan."""
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
hi = lambda o: ops.Operand(o.hi, None, o.norms, o.meta, o.n, o.d, o.ldp)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
a = ops.compact_operand(ops.prepare_operand(X))
for label, c in (("random order (events)", cent), ("column 0 dominates (no events)", None)):
    if c is None:
        c = cent.clone(); c[0] = 1.0  # all-positive data: <x, ones> beats every unit centroid
    b = ops.prepare_operand(c)
    print(label)
    print("   split  (2 products): %.3f ms" % t(lambda: ops.gemm_select(a, b, METRIC_IP, 1)))
    print("   coarse (1 product) : %.3f ms" % t(lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 1)))
