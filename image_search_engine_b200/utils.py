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

    Behaviour kept from the reference, the default ``compat=True``:
      * ``fit`` computes idf = log((n - df + 0.5) / (df + 0.5)) and stores it, but ``transform`` never
        applies it, nor the ``norm`` parameter -- the output is tf*k1 / (tf*k1 + k2*(1 - b + b*dl/avgdl));
      * ``avgdl`` is the mean document length of the batch being transformed, so a single query row
        always sees dl/avgdl == 1;
      * host inputs come back as a float64 scipy CSR matrix.
    ``compat=False`` is the opt-in corrected mode the reference's parameters promise: the tf weights are
    multiplied by the fitted idf (when ``use_idf``) and every row is then l1 / l2 normalised (``norm``).

    Host inputs are weighted as CSR: only the non-zeros travel to the device and back (O(nnz), the reference
    works on ``X.data`` the same way, utils.py:180-200).  A CUDA tensor input is weighted in place and returned
    as a tensor (GPU-resident index build).
    """

    def __init__(self, *, norm="l2", use_idf=True, k1=1, k2=1, b=0.75, compat=True):
        self.norm = norm
        self.use_idf = use_idf
        self.k1 = k1
        self.k2 = k2
        self.b = b
        self.compat = compat

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

    def _norm_code(self):
        if getattr(self, "compat", True) or self.norm is None:
            return 0
        if self.norm not in ("l1", "l2"):
            raise ValueError(f"norm must be 'l1', 'l2' or None, got {self.norm!r}")
        return 1 if self.norm == "l1" else 2

    def _idf_device(self, dev):
        if getattr(self, "compat", True) or not self.use_idf:
            return None
        return torch.from_numpy(np.ascontiguousarray(self.idf_, dtype=np.float64)).to(dev)

    def finish_device_(self, H: torch.Tensor) -> torch.Tensor:
        """Second half of the opt-in mode on an [n, k] CUDA matrix already holding the Okapi tf weights (the
        histogram kernels fuse those): idf scaling + row normalisation.  No-op with ``compat=True``."""
        norm, idf = self._norm_code(), self._idf_device(H.device)
        if norm or idf is not None:
            ops.tfidf_finish_(H, idf, norm)
        return H

    def transform(self, X, copy=True):
        if isinstance(X, torch.Tensor) and X.is_cuda:
            H = X.clone() if copy else X
            ops.okapi_tf_(H, self.k1, self.k2, self.b)
            return self.finish_device_(H)
        if sp.issparse(X):
            M = X.tocsr()
            if M is X and copy:
                M = M.copy()
            if M.dtype not in (np.float64, np.float32):
                M = M.astype(np.float64)
        else:
            dense = np.asarray(X)
            if dense.ndim != 2:
                raise ValueError("Expected a 2-D count matrix")
            M = sp.csr_matrix(dense, dtype=np.float64)
        if M.nnz == 0:
            return M
        dev = ops.require_cuda()
        norm = self._norm_code()
        idf = self._idf_device(dev)
        data = torch.from_numpy(np.ascontiguousarray(M.data, dtype=np.float64)).to(dev, non_blocking=True)
        indptr = torch.from_numpy(np.ascontiguousarray(M.indptr, dtype=np.int64)).to(dev, non_blocking=True)
        indices = None
        if idf is not None:
            indices = torch.from_numpy(np.ascontiguousarray(M.indices, dtype=np.int32)).to(dev, non_blocking=True)
        ops.okapi_csr_(indptr, indices, data, self.k1, self.k2, self.b, idf=idf, norm=norm)
        M.data[...] = data.cpu().numpy()      # float32 matrices keep their dtype (computed in float64)
        return M

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
