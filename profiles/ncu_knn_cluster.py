"""ncu driver: one seeded coarse top-32 launch on the C3 shape per multicast cluster size (ISE_CLUSTER_PAIRS 1, 2, 4)."""
import os, sys
import torch
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
b = ops.attach_sample(ops.prepare_operand(db)); a = ops.prepare_operand(q, rows=True)
seed = ops.gemm_select(a.hi_only(), b.sample, METRIC_IP, 2)[0][:, 1].contiguous()
for cl in ("1", "2", "4"):
    os.environ["ISE_CLUSTER_PAIRS"] = cl
    for _ in range(2):
        ops.gemm_select(a.hi_only(), b.hi_only(), METRIC_IP, 32, row_seed=seed)
torch.cuda.synchronize()
print("ok")
