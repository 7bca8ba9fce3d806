"""NumPy restatement of the Faiss subset that image-search-engine calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: faiss-cpu
is a third-party dependency whose sources are not under /root/reference and
whose version the reference does not pin; this file restates the published
algorithms (upstream faiss ~v1.7.4) and every function names the upstream
routine it follows plus the reference call site that needs it.

Reference call sites (all relative to /root/reference/backend):
  faiss.Kmeans(...).train / .index / .centroids / .obj   kmeans_faiss.py:29-44
  index.search(x, 1)                                     kmeans_faiss.py:49
  faiss.IndexFlatIP / IndexFlatL2 / normalize_L2 / add   utils.py:300-327
  index.search(q, n)                                     engine.py:55, siamese/test_index.py:54
  faiss.write_index / read_index                         bag_of_visual_words.py:187,194,213; engine.py:116,131
"""
from __future__ import annotations

import struct
import time

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

# faiss/utils/distances.cpp globals
distance_compute_blas_threshold = 20
distance_compute_blas_query_bs = 4096
distance_compute_blas_database_bs = 1024
distance_compute_min_k_reservoir = 100

FLT_MAX = np.float32(np.finfo(np.float32).max)


# --------------------------------------------------------------------------
# faiss/utils/random.cpp
# --------------------------------------------------------------------------
class RandomGenerator:
    """std::mt19937 seeded with (unsigned)seed  (faiss RandomGenerator)."""

    def __init__(self, seed: int):
        bg = np.random.MT19937()
        bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
        self._bg = bg

    def raw(self, n: int) -> np.ndarray:
        return self._bg.random_raw(int(n)).astype(np.uint64)

    def rand_float_batch(self, n: int) -> np.ndarray:
        # rand_float(): mt() / float(mt.max())  -- both operands as float32
        r = self.raw(n).astype(np.float32)
        return r / np.float32(4294967295.0)


def rand_perm(n: int, seed: int, prefix: int | None = None) -> np.ndarray:
    """faiss::rand_perm: Fisher-Yates with j = i + mt() % (n - i).

    ``prefix`` returns only perm[:prefix] (entries are final once passed), which
    is all Clustering::train ever reads (k or k*256 entries).
    """
    n = int(n)
    m = n if prefix is None else min(int(prefix), n)
    steps = min(m, max(n - 1, 0))
    rng = RandomGenerator(seed)
    raws = rng.raw(steps) if steps > 0 else np.zeros(0, np.uint64)
    js = (np.arange(steps, dtype=np.uint64)
          + raws % (np.uint64(n) - np.arange(steps, dtype=np.uint64))).astype(np.int64)
    if prefix is None or m * 4 >= n:
        perm = np.arange(n, dtype=np.int64)
        for i in range(steps):
            j = js[i]
            perm[i], perm[j] = perm[j], perm[i]
        return perm[:m].copy()
    # sparse variant: only displaced slots are tracked
    moved: dict[int, int] = {}
    out = np.empty(m, dtype=np.int64)
    for i in range(steps):
        j = int(js[i])
        vi = moved.get(i, i)
        vj = moved.get(j, j)
        moved[j] = vi
        out[i] = vj
    for i in range(steps, m):
        out[i] = moved.get(i, i)
    return out


# --------------------------------------------------------------------------
# faiss/utils/distances.cpp
# --------------------------------------------------------------------------
def fvec_norms_L2sqr(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.einsum("ij,ij->i", x, x, dtype=np.float32).astype(np.float32)


def normalize_L2(x: np.ndarray) -> None:
    """faiss.normalize_L2 = fvec_renorm_L2: in place, rows with zero norm untouched
    (reference: utils.py:303, engine.py:53, siamese/test_index.py:53)."""
    a = np.asarray(x)  # shares memory with np.matrix inputs
    if a.dtype != np.float32 or a.ndim != 2 or not a.flags.c_contiguous:
        raise TypeError("normalize_L2 needs a C-contiguous float32 2-D array")
    nr = fvec_norms_L2sqr(a)
    nz = nr > 0
    inv = np.ones_like(nr)
    inv[nz] = np.float32(1.0) / np.sqrt(nr[nz], dtype=np.float32)
    a *= inv[:, None]


def _select_topk(scores: np.ndarray, ids: np.ndarray, k: int, largest: bool):
    """Canonical order: best score first, ties by ascending id."""
    key = -scores if largest else scores
    if scores.shape[1] > 4 * k + 8:
        part = np.argpartition(key, k - 1, axis=1)[:, :k]
        kth = np.take_along_axis(key, part, axis=1).max(axis=1)
        out_s = np.empty((scores.shape[0], k), scores.dtype)
        out_i = np.empty((scores.shape[0], k), np.int64)
        for r in range(scores.shape[0]):
            cand = np.nonzero(key[r] <= kth[r])[0]
            order = np.lexsort((ids[r, cand], key[r, cand]))[:k]
            sel = cand[order]
            out_s[r] = scores[r, sel]
            out_i[r] = ids[r, sel]
        return out_s, out_i
    order = np.lexsort((ids, key), axis=1)[:, :k]
    return np.take_along_axis(scores, order, axis=1), np.take_along_axis(ids, order, axis=1)


def knn(x: np.ndarray, y: np.ndarray, k: int, metric: int,
        y_norms: np.ndarray | None = None, db_block: int | None = None):
    """knn_inner_product / knn_L2sqr (IndexFlat::search).

    n < 20  : exhaustive_*_seq -- direct per-pair dot / sum (x-y)^2.
    n >= 20 : exhaustive_*_blas -- sgemm cross term; L2 = |x|^2+|y|^2-2<x,y> clamped at 0.
    Results: best first (IP descending, L2 ascending squared distance); ties by
    ascending id (k==1: strict compare, lowest id wins); k > ntotal pads with
    id -1 and -FLT_MAX (IP) / +FLT_MAX (L2).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    nx, d = x.shape
    ny = y.shape[0]
    largest = metric == METRIC_INNER_PRODUCT
    pad = -FLT_MAX if largest else FLT_MAX
    D = np.full((nx, k), pad, dtype=np.float32)
    I = np.full((nx, k), -1, dtype=np.int64)
    if nx == 0 or ny == 0:
        return D, I
    blas = nx >= distance_compute_blas_threshold
    bs_x = distance_compute_blas_query_bs
    bs_y = db_block or (distance_compute_blas_database_bs if k > 1 else 16384)
    if not largest and blas:
        x_norms = fvec_norms_L2sqr(x)
        if y_norms is None:
            y_norms = fvec_norms_L2sqr(y)
    for i0 in range(0, nx, bs_x):
        i1 = min(nx, i0 + bs_x)
        xb = x[i0:i1]
        run_s = np.full((i1 - i0, k), pad, dtype=np.float32)
        run_i = np.full((i1 - i0, k), -1, dtype=np.int64)
        for j0 in range(0, ny, bs_y):
            j1 = min(ny, j0 + bs_y)
            yb = y[j0:j1]
            if blas:
                ip = xb @ yb.T
                if largest:
                    s = ip
                else:
                    s = x_norms[i0:i1, None] + y_norms[None, j0:j1] - np.float32(2.0) * ip
                    np.maximum(s, np.float32(0.0), out=s)
            else:
                if largest:
                    s = np.einsum("id,jd->ij", xb, yb, dtype=np.float32)
                else:
                    diff = xb[:, None, :] - yb[None, :, :]
                    s = np.einsum("ijd,ijd->ij", diff, diff, dtype=np.float32)
            s = s.astype(np.float32, copy=False)
            if k == 1:
                # Top1BlockResultHandler: strict compare, first occurrence wins
                a = s.argmax(axis=1) if largest else s.argmin(axis=1)
                v = s[np.arange(s.shape[0]), a]
                better = (v > run_s[:, 0]) if largest else (v < run_s[:, 0])
                run_s[better, 0] = v[better]
                run_i[better, 0] = a[better] + j0
            else:
                ids = np.broadcast_to(np.arange(j0, j1, dtype=np.int64), s.shape)
                kk = min(k, s.shape[1])
                bs_, bi_ = _select_topk(s, ids, kk, largest)
                cs = np.concatenate([run_s, bs_], axis=1)
                ci = np.concatenate([run_i, bi_], axis=1)
                # padded (-1) slots must lose every tie
                ci_key = np.where(ci < 0, np.iinfo(np.int64).max, ci)
                key = -cs if largest else cs
                order = np.lexsort((ci_key, key), axis=1)[:, :k]
                run_s = np.take_along_axis(cs, order, axis=1)
                run_i = np.take_along_axis(ci, order, axis=1)
        D[i0:i1] = run_s
        I[i0:i1] = run_i
    return D, I


# --------------------------------------------------------------------------
# faiss/IndexFlat.cpp
# --------------------------------------------------------------------------
class IndexFlat:
    def __init__(self, d: int, metric: int = METRIC_L2):
        self.d = int(d)
        self.metric_type = int(metric)
        self.metric_arg = 0.0
        self.is_trained = True
        self.verbose = False
        self._xb = np.zeros((0, self.d), dtype=np.float32)
        self._norms = None

    @property
    def ntotal(self) -> int:
        return int(self._xb.shape[0])

    def add(self, x) -> None:
        x = np.asarray(x)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError("add: expected (n, %d) array" % self.d)
        x = np.ascontiguousarray(x, dtype=np.float32)
        self._xb = np.concatenate([self._xb, x], axis=0)
        self._norms = None

    def train(self, x) -> None:  # flat indexes need no training
        pass

    def reset(self) -> None:
        self._xb = np.zeros((0, self.d), dtype=np.float32)
        self._norms = None

    def search(self, x, k: int):
        x = np.asarray(x)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError("search: expected (n, %d) array" % self.d)
        if k <= 0:
            raise AssertionError("k must be positive")
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.metric_type == METRIC_L2 and self._norms is None:
            self._norms = fvec_norms_L2sqr(self._xb)
        return knn(x, self._xb, int(k), self.metric_type, self._norms)

    def reconstruct(self, i: int) -> np.ndarray:
        return self._xb[int(i)].copy()

    def reconstruct_n(self, i0: int = 0, ni: int = -1) -> np.ndarray:
        if ni < 0:
            ni = self.ntotal - i0
        return self._xb[i0:i0 + ni].copy()

    def __repr__(self):
        name = "IndexFlatIP" if self.metric_type == METRIC_INNER_PRODUCT else "IndexFlatL2"
        return f"<oracle.faiss_shim.{name} d={self.d} ntotal={self.ntotal}>"


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int):
        super().__init__(d, METRIC_L2)


# --------------------------------------------------------------------------
# faiss/impl/ProductQuantizer.cpp + faiss/IndexIVFPQ.cpp (the "cell-probe" branch, utils.py:311-325)
# Restated from the published algorithm (faiss ~1.7.4): by_residual = true, L2, flat coarse quantizer,
# one independent k-means (Clustering defaults: niter 25, seed 1234, <= 256 points per centroid) per
# sub-quantizer, level-1 clustering with niter = 10.  Distances are evaluated with the direct look-up-table
# form (sum over sub-quantizers of |(q - c)_m - pq_m[code_m]|^2); Faiss's precomputed-table variant is the
# same quantity regrouped, i.e. equal up to FP32 rounding.  Nothing in the reference pins this path.
# --------------------------------------------------------------------------
class ProductQuantizer:
    def __init__(self, d: int, M: int, nbits: int):
        if d % M != 0:
            raise RuntimeError("The dimension of the vector (d) should be a multiple of the number of subquantizers (M)")
        self.d, self.M, self.nbits = int(d), int(M), int(nbits)
        self.dsub, self.ksub = self.d // self.M, 1 << self.nbits
        self.cp = ClusteringParameters()
        self.centroids = np.zeros((self.M, self.ksub, self.dsub), dtype=np.float32)

    def train(self, x: np.ndarray) -> None:                    # ProductQuantizer::train, Train_default
        x = np.ascontiguousarray(x, dtype=np.float32)
        for m in range(self.M):
            xs = np.ascontiguousarray(x[:, m * self.dsub:(m + 1) * self.dsub])
            clus = Clustering(self.dsub, self.ksub, self.cp)
            clus.train(xs, IndexFlatL2(self.dsub))
            self.centroids[m] = clus.centroids.reshape(self.ksub, self.dsub)

    def compute_codes(self, x: np.ndarray) -> np.ndarray:      # nearest sub-centroid, lowest index on ties
        x = np.ascontiguousarray(x, dtype=np.float32)
        codes = np.zeros((x.shape[0], self.M), dtype=np.uint8)
        for m in range(self.M):
            xs = np.ascontiguousarray(x[:, m * self.dsub:(m + 1) * self.dsub])
            codes[:, m] = knn(xs, self.centroids[m], 1, METRIC_L2)[1].ravel().astype(np.uint8)
        return codes

    def compute_distance_table(self, x: np.ndarray) -> np.ndarray:     # [M, ksub] of one vector: fvec_L2sqr_ny
        xs = np.asarray(x, dtype=np.float32).reshape(self.M, 1, self.dsub)
        diff = xs - self.centroids
        return np.einsum("mkj,mkj->mk", diff, diff, dtype=np.float32)


class IndexIVFPQ:
    def __init__(self, quantizer: IndexFlat, d: int, nlist: int, M: int, nbits_per_idx: int, metric: int = METRIC_L2):
        if metric != METRIC_L2:
            raise NotImplementedError("the reference builds IndexIVFPQ with the default L2 metric only")
        self.quantizer, self.d, self.nlist = quantizer, int(d), int(nlist)
        self.pq = ProductQuantizer(d, M, nbits_per_idx)
        self.cp = ClusteringParameters()
        self.cp.niter = 10                                      # IndexIVF level-1 clustering
        self.nprobe, self.by_residual, self.is_trained, self.metric_type = 1, True, False, METRIC_L2
        self.ids = [np.zeros(0, np.int64) for _ in range(self.nlist)]
        self.codes = [np.zeros((0, self.pq.M), np.uint8) for _ in range(self.nlist)]
        self.ntotal = 0

    def train(self, x) -> None:
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
        if not (self.quantizer.is_trained and self.quantizer.ntotal == self.nlist):     # train_q1
            clus = Clustering(self.d, self.nlist, self.cp)
            self.quantizer.reset()
            clus.train(x, self.quantizer)
        nmax = self.pq.cp.max_points_per_centroid * self.pq.ksub                         # train_residual
        if x.shape[0] > nmax:                                                            # fvecs_maybe_subsample
            x = np.ascontiguousarray(x[rand_perm(x.shape[0], self.pq.cp.seed, prefix=nmax)])
        assign = self.quantizer.search(x, 1)[1].ravel()
        self.pq.train(x - self.quantizer._xb[assign])
        self.is_trained = True

    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed")
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
        assign = self.quantizer.search(x, 1)[1].ravel()
        codes = self.pq.compute_codes(x - self.quantizer._xb[assign])
        ids = np.arange(self.ntotal, self.ntotal + x.shape[0], dtype=np.int64)
        for l in range(self.nlist):
            sel = assign == l
            self.ids[l] = np.concatenate([self.ids[l], ids[sel]])
            self.codes[l] = np.concatenate([self.codes[l], codes[sel]])
        self.ntotal += x.shape[0]

    def search(self, x, k: int):
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
        nq, nprobe = x.shape[0], min(self.nprobe, self.nlist)
        coarse = self.quantizer.search(x, nprobe)[1]
        D = np.full((nq, k), np.finfo(np.float32).max, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        M = self.pq.M
        for qi in range(nq):
            ds, ids = [], []
            for key in coarse[qi]:
                if key < 0 or self.ids[key].size == 0:
                    continue
                tab = self.pq.compute_distance_table(x[qi] - self.quantizer._xb[key])
                dis = np.zeros(self.ids[key].size, dtype=np.float32)
                for m in range(M):                               # dis0 + sum_m table[m][code_m], in this order
                    dis += tab[m][self.codes[key][:, m]]
                ds.append(dis)
                ids.append(self.ids[key])
            if not ds:
                continue
            ds, ids = np.concatenate(ds), np.concatenate(ids)
            order = np.lexsort((ids, ds))[:k]                    # canonical (distance, id) order
            D[qi, :order.size], I[qi, :order.size] = ds[order], ids[order]
        return D, I


# --------------------------------------------------------------------------
# faiss/impl/index_write.cpp / index_read.cpp (IndexFlat only)
# --------------------------------------------------------------------------
def write_index(index: IndexFlat, fname: str) -> None:
    fourcc = b"IxFI" if index.metric_type == METRIC_INNER_PRODUCT else b"IxF2"
    with open(str(fname), "wb") as f:
        f.write(fourcc)
        f.write(struct.pack("<i", index.d))
        f.write(struct.pack("<q", index.ntotal))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<B", 1 if index.is_trained else 0))
        f.write(struct.pack("<i", index.metric_type))
        if index.metric_type > 1:
            f.write(struct.pack("<f", index.metric_arg))
        f.write(struct.pack("<Q", index.ntotal * index.d))
        f.write(np.ascontiguousarray(index._xb, dtype="<f4").tobytes())


def read_index(fname: str) -> IndexFlat:
    with open(str(fname), "rb") as f:
        fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise RuntimeError("unsupported index fourcc %r" % fourcc)
        (d,) = struct.unpack("<i", f.read(4))
        (ntotal,) = struct.unpack("<q", f.read(8))
        f.read(16)
        (is_trained,) = struct.unpack("<B", f.read(1))
        (metric,) = struct.unpack("<i", f.read(4))
        if metric > 1:
            f.read(4)
        (count,) = struct.unpack("<Q", f.read(8))
        if count != ntotal * d:
            raise RuntimeError("corrupt flat index payload")
        xb = np.frombuffer(f.read(count * 4), dtype="<f4").reshape(ntotal, d)
    idx = IndexFlatIP(d) if fourcc == b"IxFI" else IndexFlatL2(d)
    idx.metric_type = metric
    idx.add(xb)
    return idx


# --------------------------------------------------------------------------
# faiss/Clustering.cpp
# --------------------------------------------------------------------------
class ClusteringParameters:
    def __init__(self):
        self.niter = 25
        self.nredo = 1
        self.verbose = False
        self.spherical = False
        self.int_centroids = False
        self.update_index = False
        self.frozen_centroids = False
        self.min_points_per_centroid = 39
        self.max_points_per_centroid = 256
        self.seed = 1234
        self.decode_block_size = 32768


EPS = 1.0 / 1024.0


def compute_centroids(d, k, x, assign, hassign, centroids):
    """FP32 sums strictly in data order per centroid, then *= 1/count (non-empty only)."""
    centroids[:] = 0
    np.add.at(hassign, assign, np.float32(1.0))
    np.add.at(centroids, assign, x)  # unbuffered, sequential in i, float32
    nz = hassign != 0
    norm = (np.float32(1.0) / hassign[nz]).astype(np.float32)
    centroids[nz] *= norm[:, None]


def split_clusters(d, k, n, hassign, centroids) -> int:
    """Void clusters take a perturbed copy of a big one; RandomGenerator(1234) fresh per call."""
    empties = np.nonzero(hassign == 0)[0]
    if empties.size == 0:
        return 0
    rng = RandomGenerator(1234)
    nsplit = 0
    fac = np.empty(d, dtype=np.float32)
    fac[0::2] = np.float32(1 + EPS)
    fac[1::2] = np.float32(1 - EPS)
    fac_j = np.empty(d, dtype=np.float32)
    fac_j[0::2] = np.float32(1 - EPS)
    fac_j[1::2] = np.float32(1 + EPS)
    denom = np.float32(n - k)
    buf = np.zeros(0, np.float32)
    pos = 0
    for ci in empties:
        cj = 0
        while True:
            if pos >= buf.size:
                buf = rng.rand_float_batch(max(1024, k))
                pos = 0
            # p = (hassign[cj] - 1.0) / (float)(n - k)   (double arithmetic, stored to float)
            p = np.float32((float(hassign[cj]) - 1.0) / float(denom))
            r = buf[pos]
            pos += 1
            if r < p:
                break
            cj = (cj + 1) % k
        centroids[ci] = centroids[cj]
        centroids[ci] *= fac
        centroids[cj] *= fac_j
        hassign[ci] = hassign[cj] / np.float32(2)
        hassign[cj] -= hassign[ci]
        nsplit += 1
    return nsplit


def imbalance_factor(n, k, assign) -> float:
    hist = np.bincount(assign, minlength=k).astype(np.float64)
    tot = hist.sum()
    uf = (hist * hist).sum()
    return float(uf * k / (tot * tot))


class Clustering:
    def __init__(self, d: int, k: int, cp: ClusteringParameters | None = None):
        self.d = int(d)
        self.k = int(k)
        cp = cp or ClusteringParameters()
        for name, val in vars(cp).items():
            setattr(self, name, val)
        self.centroids = np.zeros(0, dtype=np.float32)
        self.iteration_stats: list[dict] = []
        self.trace = None  # optional: list collecting per-iteration dumps for lock-step tests

    def post_process_centroids(self):
        c = self.centroids.reshape(-1, self.d)
        if self.spherical:
            normalize_L2(c)
        if self.int_centroids:
            np.rint(c, out=c)

    def train(self, x: np.ndarray, index: IndexFlat, weights=None):
        if weights is not None:
            raise NotImplementedError("weights are never passed by the reference")
        d, k = self.d, self.k
        x = np.ascontiguousarray(x, dtype=np.float32)
        nx = x.shape[0]
        if nx < k:
            raise RuntimeError(
                "Number of training points (%d) should be at least as large as number of clusters (%d)" % (nx, k))
        if index.d != d:
            raise RuntimeError("Index dimension %d not the same as data dimension %d" % (index.d, d))
        t0 = time.time()
        if not np.isfinite(x).all():
            raise RuntimeError("input contains NaN's or Inf's")
        if nx > k * self.max_points_per_centroid:
            nsub = k * self.max_points_per_centroid
            perm = rand_perm(nx, self.seed, prefix=nsub)
            x = np.ascontiguousarray(x[perm])
            nx = nsub
        if nx == k:
            self.centroids = x.reshape(-1).copy()
            self.iteration_stats.append(dict(obj=0.0, time=0.0, time_search=0.0, imbalance_factor=1.0, nsplit=0))
            index.reset()
            index.add(self.centroids.reshape(k, d))
            return
        lower_is_better = index.metric_type != METRIC_INNER_PRODUCT
        best_obj = np.float32(np.inf) if lower_is_better else np.float32(-np.inf)
        best_stats, best_centroids = [], None
        if self.centroids.size % d != 0:
            raise RuntimeError("size of provided input centroids not a multiple of dimension")
        n_input = self.centroids.size // d
        input_centroids = self.centroids.reshape(n_input, d).copy()
        t_search = 0.0
        for redo in range(self.nredo):
            cent = np.zeros((k, d), dtype=np.float32)
            cent[:n_input] = input_centroids[:k]
            perm = rand_perm(nx, self.seed + 1 + redo * 15486557, prefix=k)
            if n_input < k:
                cent[n_input:] = x[perm[n_input:k]]
            self.centroids = cent.reshape(-1)
            self.post_process_centroids()
            if index.ntotal != 0:
                index.reset()
            index.add(cent)
            obj = np.float32(0)
            for it in range(self.niter):
                ts = time.time()
                dis, assign = index.search(x, 1)
                dis, assign = dis.ravel(), assign.ravel()
                t_search += time.time() - ts
                # obj: float accumulator, sequential order
                obj = np.cumsum(dis, dtype=np.float32)[-1]
                if self.trace is not None:
                    self.trace.append(dict(redo=redo, it=it, centroids_in=cent.copy(),
                                           assign=assign.copy(), dis=dis.copy()))
                hassign = np.zeros(k, dtype=np.float32)
                compute_centroids(d, k, x, assign, hassign, cent)
                nsplit = split_clusters(d, k, nx, hassign, cent)
                self.iteration_stats.append(dict(
                    obj=float(obj), time=time.time() - t0, time_search=t_search,
                    imbalance_factor=imbalance_factor(nx, k, assign), nsplit=nsplit))
                self.post_process_centroids()
                if self.trace is not None:
                    self.trace[-1]["centroids_out"] = cent.copy()
                    self.trace[-1]["nsplit"] = nsplit
                index.reset()
                index.add(cent)
            if self.nredo > 1:
                if (lower_is_better and obj < best_obj) or (not lower_is_better and obj > best_obj):
                    best_centroids = cent.copy()
                    best_stats = list(self.iteration_stats)
                    best_obj = obj
                index.reset()
        if self.nredo > 1:
            self.centroids = best_centroids.reshape(-1)
            self.iteration_stats = best_stats
            index.reset()
            index.add(best_centroids)


# --------------------------------------------------------------------------
# faiss/python/extra_wrappers.py : class Kmeans
# --------------------------------------------------------------------------
class Kmeans:
    def __init__(self, d, k, **kwargs):
        self.d = int(d)
        self.k = int(k)
        self.gpu = False
        self.cp = ClusteringParameters()
        for key, v in kwargs.items():
            if key == "gpu":
                raise NotImplementedError("gpu= is not part of the CPU oracle")
            getattr(self.cp, key)  # AttributeError on unknown field, like the SWIG object
            setattr(self.cp, key, v)
        self.centroids = None
        self.obj = None
        self.iteration_stats = None
        self.index = None
        self.trace = None

    def train(self, x, weights=None, init_centroids=None):
        x = np.ascontiguousarray(x, dtype="float32")
        n, d = x.shape
        assert d == self.d
        clus = Clustering(d, self.k, self.cp)
        clus.trace = self.trace
        if init_centroids is not None:
            nc, d2 = init_centroids.shape
            assert d2 == d
            clus.centroids = np.ascontiguousarray(init_centroids, dtype=np.float32).ravel().copy()
        self.index = IndexFlatIP(d) if self.cp.spherical else IndexFlatL2(d)
        clus.train(x, self.index, weights)
        self.centroids = clus.centroids.reshape(self.k, d)
        self.iteration_stats = clus.iteration_stats
        self.obj = np.array([st["obj"] for st in clus.iteration_stats])
        return self.obj[-1] if self.obj.size > 0 else 0.0

    def assign(self, x):
        x = np.ascontiguousarray(x, dtype="float32")
        assert self.centroids is not None, "should train before assigning"
        D, I = self.index.search(x, 1)
        return D.ravel(), I.ravel()
