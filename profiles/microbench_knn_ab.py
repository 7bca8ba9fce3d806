import sys, torch
sys.path.insert(0, ".")
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
nb, d, nq = 1_000_000, 2048, 10_000
dev = ops.require_cuda()
g = torch.Generator(device=dev); g.manual_seed(3)
db = torch.empty((nb, d), device=dev)
for i in range(0, nb, 100_000):
    db[i:i + 100_000].normal_(generator=g).clamp_(min=0)
ops.normalize_l2_(db)
q = db[torch.randint(0, nb, (nq,), generator=g, device=dev)] + 0.05 * torch.randn((nq, d), generator=g, device=dev)
ops.normalize_l2_(q)
b = ops.attach_sample(ops.prepare_operand(db)); a = ops.prepare_operand(q)
hi = lambda o: ops.Operand(o.hi, None, o.norms, o.meta, o.n, o.d, o.ldp)
seed = ops.gemm_select(hi(a), b.sample, METRIC_IP, 2)[0][:, 1].contiguous()
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("coarse k=32 seeded: %.2f ms" % t(lambda: ops.gemm_select(hi(a), hi(b), METRIC_IP, 32, row_seed=seed)), flush=True)
print("search_topk verified: %.2f ms" % t(lambda: ops.search_topk(q, a, db, b, METRIC_IP, 10)), flush=True)
print("stats:", ops.last_search_stats)
