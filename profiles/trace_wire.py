import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import sift_like, C2
from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, faiss_compat, ops
from image_search_engine_b200.bag_of_visual_words import PackedDescriptions
dev = ops.require_cuda()
rng = np.random.default_rng(1)
X = sift_like(rng, C2["n_desc"], C2["d"])
cent = X[rng.choice(len(X), C2["k"], replace=False)].copy(); cent /= np.linalg.norm(cent, axis=1, keepdims=True)
gi = faiss_compat.IndexFlatIP(C2["d"]); gi.add(cent)
bovw = BOVW(None, n_clusters=C2["k"]); bovw.clusterer = FaissKMeans(C2["k"], index=gi)
ok = OkapiTransformer()
offsets = np.arange(0, C2["n_desc"] + 1, C2["per_img"], dtype=np.int64)
packed = PackedDescriptions(torch.from_numpy(X), offsets).pin()
os.environ["ISE_NARROW_PINNED"] = "1"
for _ in range(4): bovw.transform_csr(packed, okapi=ok, n_chunks=16, copy=False)
os.environ["ISE_TRACE_WIRE"] = "1"
for _ in range(2):
    t=time.perf_counter(); bovw.transform_csr(packed, okapi=ok, n_chunks=16, copy=False); print("total ms", (time.perf_counter()-t)*1e3)
