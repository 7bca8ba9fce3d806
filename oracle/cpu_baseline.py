"""CPU restatement of the reference's hot-path LOOPS on top of the Faiss shim, for bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Each function restates the control flow of the reference file:line it cites, so the timing includes
the per-image Python loop the reference really pays (bag_of_visual_words.py:98-106).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import faiss_shim as faiss


def transform_words(index, X):
    """FaissKMeans.transform (kmeans_faiss.py:46-50)."""
    _, I = index.search(X.astype(np.float32), 1)
    return I


def visual_word_histograms(index, images_descriptions, n_clusters):
    """create_visual_word_histogram (bag_of_visual_words.py:98-106): one search + one np.histogram per image."""
    H = np.zeros((len(images_descriptions), n_clusters))
    for i, X in enumerate(images_descriptions):
        idx = transform_words(index, X)
        values, _ = np.histogram(idx, bins=n_clusters)
        H[i] = values
    return H


def okapi_transform(X, k1=1, k2=1, b=0.75):
    """OkapiTransformer.transform (utils.py:153-202) on a dense count matrix -> CSR float64."""
    X = sp.csr_matrix(X, dtype=np.float64)
    dl = X.sum(axis=1)
    sz = X.indptr[1:] - X.indptr[0:-1]
    rep = np.repeat(np.asarray(dl), sz)
    avgdl = np.mean(dl)
    X.data *= k1
    X.data /= X.data + k2 * (1 - b + b * (rep / avgdl))
    return X


def codebook_index(centroids):
    idx = faiss.IndexFlatIP(centroids.shape[1])
    idx.add(np.ascontiguousarray(centroids, dtype=np.float32))
    return idx


def assign_histogram_step(index, images_descriptions, n_clusters):
    """One hot-path step as the reference runs it: quantise + histogram + Okapi tf."""
    H = visual_word_histograms(index, images_descriptions, n_clusters)
    return okapi_transform(H)


def flat_search(db, queries, k, metric="ip"):
    """create_search_index + index.search (utils.py:293-330, engine.py:55)."""
    idx = faiss.IndexFlatIP(db.shape[1]) if metric == "ip" else faiss.IndexFlatL2(db.shape[1])
    idx.add(db)
    return idx.search(queries, k)
