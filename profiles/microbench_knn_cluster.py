"""A/B of the multicast cluster size (ISE_CLUSTER_PAIRS = pairs per cluster) on the C3 coarse pass, with a result check."""
import os, sys
import torch
sys.path.insert(0, ".")
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
nb, d, nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 2048, 10_000
dev = ops.require_cuda()
g = torch.Generator(device=dev); g.manual_seed(3)
db = torch.empty((nb, d), device=dev)
for i in range(0, nb, 100_000):
    db[i:i + 100_000].normal_(generator=g).clamp_(min=0)
ops.normalize_l2_(db)
q = db[torch.randint(0, nb, (nq,), generator=g, device=dev)] + 0.05 * torch.randn((nq, d), generator=g, device=dev)
ops.normalize_l2_(q)
b = ops.attach_sample(ops.prepare_operand(db)); a = ops.prepare_operand(q, rows=True)
seed = ops.gemm_select(a.hi_only(), b.sample, METRIC_IP, 2)[0][:, 1].contiguous()
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for rep in range(2):
    for cl in ("1", "2", "mt2"):
        os.environ["ISE_CLUSTER_PAIRS"] = "1" if cl == "mt2" else cl
        os.environ["ISE_MT2_TOPK"] = "1" if cl == "mt2" else "0"
        v, i = ops.gemm_select(a.hi_only(), b.hi_only(), METRIC_IP, 32, row_seed=seed)
        torch.cuda.synchronize()
        if ref is None:
            ref = (v.clone(), i.clone())
        same = bool(torch.equal(i, ref[1]) and torch.equal(v, ref[0]))
        ms = t(lambda: ops.gemm_select(a.hi_only(), b.hi_only(), METRIC_IP, 32, row_seed=seed))
        D, I = ops.search_topk(q, a, db, b, METRIC_IP, 10)
        ms_s = t(lambda: ops.search_topk(q, a, db, b, METRIC_IP, 10), 3)
        print(f"pairs per cluster {cl}: coarse seeded top-32 {ms:7.2f} ms  identical to CL=1: {same}   full verified search {ms_s:7.2f} ms  {ops.last_search_stats}", flush=True)
