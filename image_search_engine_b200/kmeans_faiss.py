"""Drop-in for backend/kmeans_faiss.py (FaissKMeans :5-50) with the arithmetic on the B200.

Contract kept from the reference:
  * constructor stores ``n_clusters, n_init, max_iter, init_centroids, index`` and nothing else
    (sklearn.clone / GridSearchCV rely on that, bag_of_visual_words.py:161-169);
  * ``fit`` trains SPHERICAL k-means with seed 42, ``niter=max_iter``, ``nredo=n_init`` and exposes
    ``kmeans``, ``index``, ``cluster_centers_`` (k, d) float32 and ``inertia_ = obj[-1]``
    (kmeans_faiss.py:29-44) -- note the objective is a sum of inner products, not an SSE;
  * ``transform`` returns the nearest visual word of every row as an int64 (n, 1) array
    (kmeans_faiss.py:46-50), i.e. ``index.search(X, 1)[1]``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import faiss_compat as faiss

KMEANS_SEED = 42  # kmeans_faiss.py:30


class FaissKMeans:
    def __init__(self, n_clusters=8, n_init=3, max_iter=25, init_centroids=None, index=None):
        self.n_clusters = n_clusters
        self.n_init = n_init
        self.max_iter = max_iter
        self.init_centroids = init_centroids
        self.index = index

    def fit(self, X, y=None) -> None:
        d = int(X.shape[1])
        km = faiss.Kmeans(d=d, k=int(self.n_clusters), seed=KMEANS_SEED, niter=self.max_iter,
                          nredo=self.n_init, spherical=True, verbose=False)
        km.train(_as_descriptor_matrix(X), init_centroids=self.init_centroids)
        self.kmeans = km
        self.index = km.index
        self.cluster_centers_ = km.centroids
        self.inertia_ = km.obj[-1]

    def transform(self, X):
        if self.index is None:
            raise AttributeError("FaissKMeans has no index: call fit() or pass index=")
        _, words = self.index.search(_as_descriptor_matrix(X), 1)
        return words

    def transform_device(self, X: torch.Tensor) -> torch.Tensor:
        """Batched variant for GPU-resident descriptors: int64 [n] CUDA tensor, no host round trip."""
        if X.dim() != 2 or X.shape[1] != self.index.d:
            raise AssertionError(f"expected (n, {self.index.d}) descriptors")
        _, words = self.index._search_device(X, 1, need_distances=False)
        return words.reshape(-1)


def _as_descriptor_matrix(X):
    """The reference does ``X.astype(np.float32)`` on the host; uint8 descriptors (ORB/BRISK) are
    instead shipped as bytes and widened on the device -- same values, a quarter of the PCIe bytes."""
    if isinstance(X, torch.Tensor):
        return X
    X = np.asarray(X)
    if X.dtype == np.uint8 or X.dtype == np.float32:
        return X
    return X.astype(np.float32)
