"""Host-side cost of one chunk of BOVW.transform_csr (cProfile of the calling thread)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import sift_like, C2
from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, faiss_compat, ops
dev = ops.require_cuda()
rng = np.random.default_rng(1)
X = sift_like(rng, 125_000, C2["d"])
cent = X[rng.choice(len(X), C2["k"], replace=False)].copy()
cent /= np.linalg.norm(cent, axis=1, keepdims=True)
gi = faiss_compat.IndexFlatIP(C2["d"]); gi.add(cent)
km = FaissKMeans(C2["k"], index=gi)
xd = torch.from_numpy(X).to(dev)
for _ in range(5):
    km.transform_device(xd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    km.transform_device(xd)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per transform_device call: {(t1 - t0) / 50 * 1e6:.0f} us (enqueue only); with device drain {(t2 - t0) / 50 * 1e6:.0f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    km.transform_device(xd)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
