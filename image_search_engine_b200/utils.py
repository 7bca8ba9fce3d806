"""Drop-in for the retrieval parts of backend/utils.py: OkapiTransformer (:100-219),
create_search_index (:293-330) and chunkIt (:29-41), with the arithmetic in libise kernels.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from sklearn.base import BaseEstimator, TransformerMixin

from . import faiss_compat as faiss
from . import ops


def chunkIt(seq, num):
    """Splits ``seq`` into ~``num`` consecutive slices with the reference's boundaries
    (utils.py:29-41: cut points are int() of a float running sum of len/num)."""
    n = len(seq)
    width = n / float(num)
    pieces, cursor = [], 0.0
    while cursor < n:
        nxt = cursor + width
        pieces.append(seq[int(cursor):int(nxt)])
        cursor = nxt
    return pieces


class OkapiTransformer(TransformerMixin, BaseEstimator):
    """BM25/Okapi term-frequency weighting of a count matrix.

    Behaviour kept from the reference:
      * ``fit`` computes idf = log((n - df + 0.5) / (df + 0.5)) and stores it, but ``transform`` never
        applies it, nor the ``norm`` parameter -- the output is tf*k1 / (tf*k1 + k2*(1 - b + b*dl/avgdl));
      * ``avgdl`` is the mean document length of the batch being transformed, so a single query row
        always sees dl/avgdl == 1;
      * host inputs come back as a float64 scipy CSR matrix.
    A CUDA tensor input is weighted in place and returned as a tensor (GPU-resident index build).
    """

    def __init__(self, *, norm="l2", use_idf=True, k1=1, k2=1, b=0.75):
        self.norm = norm
        self.use_idf = use_idf
        self.k1 = k1
        self.k2 = k2
        self.b = b

    def fit(self, X, y=None):
        if self.use_idf:
            if isinstance(X, torch.Tensor):
                n_samples = X.shape[0]
                df = torch.count_nonzero(X, dim=0).to(torch.float64).cpu().numpy()
            else:
                Xs = X if sp.issparse(X) else np.asarray(X)
                n_samples = Xs.shape[0]
                df = (np.asarray((Xs != 0).sum(axis=0)).ravel() if sp.issparse(Xs)
                      else np.count_nonzero(Xs, axis=0)).astype(np.float64)
            idf = np.log((n_samples - df + 0.5) / (df + 0.5))
            self._idf_diag = sp.diags(idf, offsets=0, shape=(df.shape[0], df.shape[0]), format="csr",
                                      dtype=np.float64)
        return self

    def transform(self, X, copy=True):
        if isinstance(X, torch.Tensor) and X.is_cuda:
            H = X.clone() if copy else X
            return ops.okapi_tf_(H, self.k1, self.k2, self.b)
        dense = X.toarray() if sp.issparse(X) else np.asarray(X)
        if dense.ndim != 2:
            raise ValueError("Expected a 2-D count matrix")
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        dev = ops.require_cuda()
        H = torch.from_numpy(dense).to(dev)
        ops.okapi_tf_(H, self.k1, self.k2, self.b)
        return sp.csr_matrix(H.cpu().numpy())

    def __sklearn_is_fitted__(self):
        # stateless at transform time; lets sklearn >= 1.4 Pipeline.transform run
        return True

    @property
    def idf_(self):
        return np.ravel(self._idf_diag.sum(axis=0))

    @idf_.setter
    def idf_(self, value):
        value = np.asarray(value, dtype=np.float64)
        n = value.shape[0]
        self._idf_diag = sp.spdiags(value, diags=0, m=n, n=n, format="csr")


def create_search_index(data_array, index_type="cosine"):
    """Builds the HBM-resident flat index.

    "cosine" -> inner-product index over L2-normalised rows; like the reference (utils.py:302-303)
    the caller's array is normalised IN PLACE.  "l2" -> squared-L2 index.  "cell-probe" -> the
    reference's IndexIVFPQ configuration (approximate; off every default path of the reference).
    """
    num_features = int(data_array.shape[1])
    if index_type == "cell-probe":
        # utils.py:311-325: 8 coarse centroids, 16 sub-quantizers of 8 bits, 5 probed lists, trained on the data itself
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(num_features), num_features, 8, 16, 8)
        index.nprobe = 5
        index.train(data_array)
        index.add(data_array)
        print(f"There are {index.ntotal} images in the search index.")
        return index
    if index_type not in ("cosine", "l2"):
        raise ValueError(f"unknown index_type {index_type!r}")
    if isinstance(data_array, torch.Tensor) and data_array.is_cuda:
        rows = data_array
        if index_type == "cosine":
            faiss.normalize_L2(rows)
    else:
        host = np.asarray(data_array)                       # np.matrix shares its buffer (quirk Q4)
        if index_type == "cosine" and (host.dtype != np.float32 or not host.flags.c_contiguous):
            raise TypeError("normalize_L2 needs a C-contiguous float32 2-D array")
        rows = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).to(ops.require_cuda())
        if index_type == "cosine":                          # one upload: normalise on the device, then
            faiss.normalize_L2(rows)                        # mirror the result into the caller's array
            host[...] = rows.cpu().numpy()
    index = faiss.IndexFlatIP(num_features) if index_type == "cosine" else faiss.IndexFlatL2(num_features)
    index.add(rows)
    print(f"There are {index.ntotal} images in the search index.")
    return index


# ------------------------------------------------------------------------------------------
# cluster-quality scorer for the optional cluster-count grid search (utils.py:235-290)
# ------------------------------------------------------------------------------------------
rs = np.random.RandomState(42)          # module-level stream, like the reference (utils.py:26)
CLUSTER_EVAL_SAMPLE_SIZE = 2000         # config.py:97-100
CLUSTER_EVAL_N_SAMPLES = 10


def calc_sampled_cluster_score(estimator, X, y=None):
    """Drop-in for utils.calc_sampled_cluster_score: minus the mean Davies-Bouldin score of
    ``CLUSTER_EVAL_N_SAMPLES`` random samples of ``CLUSTER_EVAL_SAMPLE_SIZE`` descriptors, labelled by the fitted
    codebook.  The part that scales with the data set -- assigning EVERY cached descriptor to its visual word
    (utils.py:274-275) -- is one fused assign launch; the 2000-row Davies-Bouldin evaluations stay scikit-learn's.
    Like the reference it scores ``estimator.named_steps["bovw"].descriptions`` and ignores X."""
    from sklearn.metrics import davies_bouldin_score
    from .bag_of_visual_words import pack_descriptions
    bovw = estimator.named_steps["bovw"]
    mat, _ = pack_descriptions(bovw.descriptions)
    labels_ = bovw.clusterer.transform(mat).ravel()
    all_descriptions = mat.cpu().numpy() if isinstance(mat, torch.Tensor) else np.asarray(mat)
    dataset_size = all_descriptions.shape[0]
    sample_size = int(min(dataset_size, CLUSTER_EVAL_SAMPLE_SIZE))
    scores = []
    for _ in range(CLUSTER_EVAL_N_SAMPLES):
        sample_idxs = rs.choice(dataset_size, size=sample_size, replace=False)
        scores.append(davies_bouldin_score(all_descriptions[sample_idxs], labels_[sample_idxs]))
    return -1 * np.mean(scores)
