"""Micro-benchmark of gemm_select variants on the C3 shape (used to separate main-loop from epilogue cost)."""
import sys, time
import torch
sys.path.insert(0, ".")
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
nq = 10_000
dev = ops.require_cuda()
g = torch.Generator(device=dev); g.manual_seed(3)
db = torch.empty((nb, d), device=dev)
for i in range(0, nb, 100_000):
    db[i:i + 100_000].normal_(generator=g)
    if d == 2048:
        db[i:i + 100_000].clamp_(min=0)
ops.normalize_l2_(db)
q = db[torch.randint(0, nb, (nq,), generator=g, device=dev)] + 0.05 * torch.randn((nq, d), generator=g, device=dev)
ops.normalize_l2_(q)
b = ops.prepare_operand(db); a = ops.prepare_operand(q)
hi = lambda o: ops.Operand(o.hi, None, o.norms, o.meta, o.n, o.d, o.ldp)

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ops.attach_sample(b)
seed = ops.gemm_select(hi(a), b.sample, METRIC_IP, 2)[0][:, 1].contiguous()

def full_verified():
    return ops.search_topk(q, a, db, b, METRIC_IP, K)

for name, fn in [
    ("sample prepass k=2     ", lambda: ops.gemm_select(hi(a), b.sample, METRIC_IP, 2)),
    ("coarse k=32 seeded     ", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 32, row_seed=seed)),
    ("search_topk verified   ", full_verified),
    ("search_topk split      ", lambda: ops.search_topk(q, a, db, b, METRIC_IP, K, precision="split")),
    ("split  k=1 ", lambda: ops.gemm_select(a, b, METRIC_IP, 1)),
    ("coarse k=1 ", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 1)),
    ("coarse k=32", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 32)),
    ("coarse k=128", lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 128)),
]:
    ms = t(fn)
    print(f"{name}: {ms:8.2f} ms  {2.0 * nq * nb * d / ms / 1e9:8.1f} algorithmic TFLOP/s", flush=True)
full_verified()
if b.sample_dense is not None and K > 16:
    rc = None
print("fallback rows in last verified search:", ops.last_search_stats)
cv, ci = ops.gemm_select(hi(a), hi(b), METRIC_IP, 32, row_seed=seed)
print("mean candidates per query kept by the seeded coarse pass:", float((ci >= 0).sum(1).float().mean()))
