"""e2e of BOVW.transform_csr(pinned float32 PackedDescriptions) at the C2 shape: wire policy x number of chunks."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import sift_like, C2
from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, faiss_compat, ops
from image_search_engine_b200.bag_of_visual_words import PackedDescriptions
dev = ops.require_cuda()
rng = np.random.default_rng(1)
X = sift_like(rng, C2["n_desc"], C2["d"])
cent = X[rng.choice(len(X), C2["k"], replace=False)].copy()
cent /= np.linalg.norm(cent, axis=1, keepdims=True)
gi = faiss_compat.IndexFlatIP(C2["d"]); gi.add(cent)
bovw = BOVW(None, n_clusters=C2["k"]); bovw.clusterer = FaissKMeans(C2["k"], index=gi)
ok = OkapiTransformer()
offsets = np.arange(0, C2["n_desc"] + 1, C2["per_img"], dtype=np.int64)
packed = PackedDescriptions(torch.from_numpy(X), offsets).pin()
for policy in ("0", "1"):
    os.environ["ISE_NARROW_PINNED"] = policy
    for nc in (4, 8, 12, 16, 24, 32):
        for _ in range(3):
            bovw.transform_csr(packed, okapi=ok, n_chunks=nc, copy=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(8):
            bovw.transform_csr(packed, okapi=ok, n_chunks=nc, copy=False)
        ms = (time.perf_counter() - t0) * 1e3 / 8
        print(f"narrow={policy} n_chunks={nc:2d}: {ms:6.2f} ms  wire: {bovw._last_transfer['wire']}  {bovw._last_transfer['h2d_bytes'] / 1e6:.0f} MB", flush=True)
