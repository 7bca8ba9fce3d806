"""Times the host->host BoVW transform (pipelined vs one-shot) on the C2 shape."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from bench import sift_like, C2
from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, ops
from image_search_engine_b200.bag_of_visual_words import PackedDescriptions

dev = ops.require_cuda()
X = sift_like(np.random.default_rng(2), C2["n_desc"], C2["d"])
offsets = np.arange(0, C2["n_desc"] + 1, C2["per_img"], dtype=np.int64)
packed = PackedDescriptions(X, offsets).pin()
km = FaissKMeans(C2["k"], n_init=1, max_iter=2); km.fit(packed.matrix.to(dev))
bovw = BOVW(None, n_clusters=C2["k"]); bovw.clusterer = km
ok = OkapiTransformer()
out = torch.empty((C2["n_img"], C2["k"]), dtype=torch.float64, pin_memory=True)

def t(fn, n=4):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def oneshot():
    H = bovw.histograms_device(packed, okapi=ok)
    out.copy_(H, non_blocking=True); torch.cuda.current_stream().synchronize()

print("one shot              : %.2f ms" % t(oneshot))
for nc in (1, 2, 4, 8, 16):
    print("pipelined n_chunks=%2d : %.2f ms" % (nc, t(lambda: bovw.histograms_host(packed, out, okapi=ok, n_chunks=nc))))
xd = packed.matrix.to(dev)
print("H2D only              : %.2f ms" % t(lambda: packed.matrix.to(dev, non_blocking=True)))
H = bovw.histograms_device(PackedDescriptions(xd, offsets), okapi=ok)
print("D2H only              : %.2f ms" % t(lambda: out.copy_(H, non_blocking=True)))
print("device only           : %.2f ms" % t(lambda: bovw.histograms_device(PackedDescriptions(xd, offsets), okapi=ok)))
for nc in (4, 8, 16):
    print("transform_csr n_chunks=%2d : %.2f ms" % (nc, t(lambda: bovw.transform_csr(packed, okapi=ok, n_chunks=nc, copy=False))))
print("csr device only       : %.2f ms" % t(lambda: bovw.csr_device(PackedDescriptions(xd, offsets), okapi=ok)))
