"""A/B on one C5 shard (1.25 M x 512, 10 k queries, top-100, collect mode): single CTAs vs CTA pairs vs multicast clusters."""
import os, sys
import torch
sys.path.insert(0, ".")
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
nb, d, nq, k = 1_250_000, 512, 10_000, 100
dev = ops.require_cuda()
g = torch.Generator(device=dev); g.manual_seed(5)
db = torch.empty((nb, d), device=dev)
for i in range(0, nb, 250_000):
    db[i:i + 250_000].normal_(generator=g)
ops.normalize_l2_(db)
q = db[torch.randint(0, nb, (nq,), generator=g, device=dev)] + 0.05 * torch.randn((nq, d), generator=g, device=dev)
ops.normalize_l2_(q)
b = ops.attach_sample(ops.prepare_operand(db)); a = ops.prepare_operand(q, rows=True)
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for rep in range(2):
    for cg2, cl in (("0", "1"), ("1", "1"), ("1", "2"), ("1", "4")):
        os.environ["ISE_CG2_COARSE"] = cg2; os.environ["ISE_CLUSTER_PAIRS"] = cl
        D, I = ops.search_topk(q, a, db, b, METRIC_IP, k)
        torch.cuda.synchronize()
        if ref is None: ref = I.clone()
        ms = t(lambda: ops.search_topk(q, a, db, b, METRIC_IP, k))
        print(f"pairs={cg2} pairs-per-cluster={cl}: full search {ms:7.2f} ms  same ids: {bool(torch.equal(I, ref))}  {ops.last_search_stats}", flush=True)
