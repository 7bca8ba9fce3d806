"""A/B of the top-1 assign at d = 128: split products (fused), verified pipeline (fused / prepared planes).
usage: python profiles/microbench_assign_verified.py [n_centroids ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import sift_like
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP, METRIC_L2

dev = ops.require_cuda()
rng = np.random.default_rng(0)
m, d = 1_000_000, 128
x = sift_like(rng, m, d)
xd = torch.from_numpy(x).to(dev)
ks = [int(a) for a in sys.argv[1:]] or [4096, 65536]


def timed(fn, reps=5):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for k in ks:
    # a trained-looking codebook: means of random groups of descriptors, unit norm
    c = x[rng.choice(m, k, replace=False)].copy()
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    cd = torch.from_numpy(c).to(dev)
    b = ops.prepare_operand(cd)
    a = ops.prepare_operand(xd, rows=True)
    for metric, name in ((METRIC_IP, "IP"), (METRIC_L2, "L2")):
        t_split, (v0, i0, _) = timed(lambda: ops.assign_fused(xd, b, metric, verified=False))
        t_ver, (v1, i1, _) = timed(lambda: ops.assign_fused(xd, b, metric, verified=True))
        st = ops.search_stats()
        t_prep, (v2, i2) = timed(lambda: ops.assign_verified(a, b, metric))
        t_gs, (v3, i3) = timed(lambda: ops.gemm_select(a, b, metric, 1))
        print(f"k={k} {name}: fused split {t_split:.3f} ms | fused verified {t_ver:.3f} ms ({st['fallback_rows']} rows re-run, "
              f"overflow {st.get('overflow')}) ids equal {bool(torch.equal(i0, i1))} | prepared verified {t_prep:.3f} ms ids equal "
              f"{bool(torch.equal(i0, i2))} | prepared split {t_gs:.3f} ms ids equal {bool(torch.equal(i0, i3))}", flush=True)
