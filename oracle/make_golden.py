"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python (imported unmodified from
/root/reference/backend by oracle/refload.py) on top of oracle/faiss_shim.py.

Run in the build container (where /root/reference exists):  python -m oracle.make_golden
The fixtures travel to the GPU box; nothing there reads /root/reference.

What each fixture pins
  bovw_c1mini.npz  C1-shaped mini pipeline: reference FaissKMeans.fit / .transform, reference
                   BOVW.transform (np.histogram loop), reference OkapiTransformer, reference
                   create_search_index("cosine" and "l2") + index.search, shim write_index bytes.
  kats.npz         known-answer cases from SURVEY section 4 (histogram quirk Q1, ties, k > ntotal).
  cluster_score.npz  reference calc_sampled_cluster_score (sampled Davies-Bouldin) on the C1-mini codebook.
The Faiss arithmetic itself comes from the shim (PARITY UNPINNED, see oracle/__init__.py); everything
else in these files is produced by reference code.
"""
from __future__ import annotations

import io
import os
import tempfile
from pathlib import Path

import numpy as np

from oracle import refload

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def c1_mini(ref, seed=1):
    rng = np.random.default_rng(seed)
    n_img, d, k = 40, 32, 32
    sizes = np.clip(np.rint(rng.normal(120, 30, n_img)), 20, 300).astype(np.int64)
    sizes[3] = 1  # single-descriptor image -> np.histogram's degenerate [v-.5, v+.5] range
    descs = [rng.integers(0, 256, size=(int(s), d), dtype=np.uint8) for s in sizes]
    X = np.concatenate(descs, axis=0)
    offsets = np.zeros(n_img + 1, np.int64)
    np.cumsum(sizes, out=offsets[1:])

    km = ref.kmeans_faiss.FaissKMeans(k, n_init=2, max_iter=4)
    km.fit(X)                                     # reference kmeans_faiss.py:25-44
    words = km.transform(X)                       # :46-50  -> (N, 1) int64

    bovw = ref.bag_of_visual_words.BOVW(describer=None, n_clusters=k)
    bovw.descriptions = descs
    bovw.clusterer = km
    hist = bovw.transform(None)                   # bag_of_visual_words.py:76-116

    okapi = ref.utils.OkapiTransformer()
    okapi.fit(hist)
    tf = okapi.transform(hist)                    # utils.py:153-202 (CSR float64)
    feats = np.asarray(tf.todense()).astype(np.float32)

    cos_in = feats.copy()
    idx_cos = ref.utils.create_search_index(cos_in, "cosine")   # normalises cos_in in place
    q = feats[:25].copy()
    D_cos, I_cos = idx_cos.search(q, 10)
    idx_l2 = ref.utils.create_search_index(feats.copy(), "l2")
    D_l2, I_l2 = idx_l2.search(q, 10)
    D_l2_1, I_l2_1 = idx_l2.search(q[:1], 10)     # nq < 20: Faiss's direct (no expansion) path
    D_cos_1, I_cos_1 = idx_cos.search(q[:1], 50)  # k > ntotal -> -1 padding

    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "codebook.faiss")
        ref.faiss.write_index(km.index, p)
        codebook_file = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)

    return dict(
        X=X, offsets=offsets, k=np.int64(k), centroids=km.cluster_centers_, obj=np.asarray(km.kmeans.obj),
        inertia=np.float64(km.inertia_), words=words, hist=hist, okapi=np.asarray(tf.todense()),
        idf=okapi.idf_, feats=feats, cos_db=cos_in, q=q, D_cos=D_cos, I_cos=I_cos, D_l2=D_l2, I_l2=I_l2,
        D_l2_1=D_l2_1, I_l2_1=I_l2_1, D_cos_1=D_cos_1, I_cos_1=I_cos_1, codebook_file=codebook_file,
        chunk_bounds=np.array([len(c) for c in ref.utils.chunkIt(list(range(37)), 5)]),
    )


def cluster_score(ref, seed=1):
    """Reference calc_sampled_cluster_score (utils.py:235-290) on the C1-mini pipeline: 10 x davies_bouldin_score
    over 2000-row samples drawn by the module's RandomState(42), applied to the labels of FaissKMeans.transform."""
    import types
    rng = np.random.default_rng(seed)
    n_img, d, k = 40, 32, 32
    sizes = np.clip(np.rint(rng.normal(120, 30, n_img)), 20, 300).astype(np.int64)
    sizes[3] = 1
    descs = [rng.integers(0, 256, size=(int(s), d), dtype=np.uint8) for s in sizes]
    km = ref.kmeans_faiss.FaissKMeans(k, n_init=2, max_iter=4)
    km.fit(np.concatenate(descs, axis=0))
    bovw = ref.bag_of_visual_words.BOVW(describer=None, n_clusters=k)
    bovw.descriptions, bovw.clusterer = descs, km
    est = types.SimpleNamespace(named_steps={"bovw": bovw})
    ref.utils.rs = np.random.RandomState(42)               # the state a fresh import of utils.py starts from
    s1 = ref.utils.calc_sampled_cluster_score(est, None)
    s2 = ref.utils.calc_sampled_cluster_score(est, None)   # the module-level stream keeps advancing
    return dict(score_first_call=np.float64(s1), score_second_call=np.float64(s2), centroids=km.cluster_centers_)


def kats(ref):
    out = {}
    idx = np.array([3, 3, 7, 190, 100, 100], dtype=np.int64).reshape(-1, 1)
    out["hist_q1_in"] = idx
    out["hist_q1_numpy"] = np.histogram(idx, bins=200)[0]
    out["hist_q1_bincount"] = np.bincount(idx.ravel(), minlength=200)
    # 4 points / 2 centroids, exact tie -> lowest id
    c = np.array([[1, 0], [1, 0], [0, 1]], dtype=np.float32)
    x = np.array([[2, 0], [0, 3], [1, 1], [5, 0]] * 6, dtype=np.float32)  # 24 rows: BLAS path
    ip = ref.faiss.IndexFlatIP(2)
    ip.add(c)
    out["tie_c"], out["tie_x"] = c, x
    out["tie_D"], out["tie_I"] = ip.search(x, 1)
    out["tie_D3"], out["tie_I3"] = ip.search(x, 3)
    out["tie_D5"], out["tie_I5"] = ip.search(x, 5)     # k > ntotal
    # Okapi on a tiny matrix with known dl/avgdl
    H = np.array([[2, 0, 1, 0], [0, 4, 0, 0], [1, 1, 1, 1]], dtype=np.float64)
    out["okapi_in"] = H
    out["okapi_out"] = np.asarray(ref.utils.OkapiTransformer().fit(H).transform(H).todense())
    out["okapi_single_row"] = np.asarray(ref.utils.OkapiTransformer().transform(H[:1]).todense())
    n = np.array([[3, 4], [0, 0], [1, 0]], dtype=np.float32)
    ref.faiss.normalize_L2(n)
    out["normalize_out"] = n
    return out


def main():
    ref = refload.load(n_clusters=32)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "bovw_c1mini.npz", **c1_mini(ref))
    np.savez_compressed(OUT / "kats.npz", **kats(ref))
    np.savez_compressed(OUT / "cluster_score.npz", **cluster_score(ref))
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
