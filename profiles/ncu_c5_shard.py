"""One verified-collect search of a C5 shard (1.25 M x 512, 10 k queries, top-100) for an ncu launch list."""
import sys
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
b = ops.attach_sample(ops.prepare_operand(db))
for _ in range(3):
    a = ops.prepare_operand(q, rows=True)
    D, I = ops.search_topk(q, a, db, b, METRIC_IP, k)
torch.cuda.synchronize()
print(ops.search_stats())
