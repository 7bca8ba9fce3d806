"""GPU tests of the round-2 boundary work: O(nnz) Okapi on CSR input, the opt-in corrected tf-idf mode, k beyond
the fused-selection limit, thread safety of the cached host pipelines, and the sharded index's small-batch / empty
shard handling.  All through the public Python surface, i.e. through the C ABI."""
import threading
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests._util import assert_topk_parity, sift_like, unit_rows

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _okapi_ref(H, k1=1.0, k2=1.0, b=0.75):
    """utils.py:180-200 restated on a dense float64 matrix (zeros stay zero)."""
    H = np.asarray(H, dtype=np.float64)
    dl = H.sum(axis=1, keepdims=True)
    t = H * k1
    with np.errstate(invalid="ignore", divide="ignore"):
        out = t / (t + k2 * (1 - b + b * (dl / dl.mean())))
    return np.where(H != 0, out, 0.0)


def test_okapi_sparse_in_sparse_out():
    from image_search_engine_b200 import OkapiTransformer
    rng = np.random.default_rng(7)
    H = np.zeros((300, 1000))
    for i in range(300):
        w = rng.integers(0, 1000, rng.integers(1, 120))
        np.add.at(H[i], w, 1.0)
    H[17] = 0                                            # an empty document
    want = _okapi_ref(H)
    X = sp.csr_matrix(H)
    T = OkapiTransformer().transform(X)                  # copy=True: the input is left alone
    assert sp.issparse(T) and T.format == "csr" and T.dtype == np.float64
    assert np.array_equal(X.toarray(), H)
    assert np.array_equal(T.indices, X.indices) and np.array_equal(T.indptr, X.indptr)
    np.testing.assert_allclose(T.toarray(), want, rtol=1e-15, atol=0)
    T2 = OkapiTransformer().transform(X, copy=False)     # in place on the caller's matrix, like the reference
    assert T2 is X
    np.testing.assert_allclose(X.toarray(), want, rtol=1e-15, atol=0)
    # other parameters, CSC input, a single row (dl / avgdl == 1)
    T3 = OkapiTransformer(k1=1.5, k2=0.5, b=0.3).transform(sp.csc_matrix(H))
    np.testing.assert_allclose(T3.toarray(), _okapi_ref(H, 1.5, 0.5, 0.3), rtol=1e-15, atol=0)
    T4 = OkapiTransformer().transform(sp.csr_matrix(H[:1]))
    np.testing.assert_allclose(T4.toarray(), H[:1] / (H[:1] + 1.0), rtol=1e-15, atol=0)
    empty = OkapiTransformer().transform(sp.csr_matrix((5, 8)))
    assert empty.shape == (5, 8) and empty.nnz == 0


def test_okapi_opt_in_idf_and_norm_known_answer():
    """compat=False: tf weights x fitted idf, then row normalisation -- checked against the closed form on a tiny
    hand-made matrix and against the dense device path (GPU-resident index build)."""
    from image_search_engine_b200 import OkapiTransformer, ops
    H = np.array([[2, 0, 1, 0],
                  [0, 3, 1, 0],
                  [1, 1, 0, 0],
                  [0, 0, 4, 0],
                  [1, 0, 0, 5]], dtype=np.float64)
    n = H.shape[0]
    df = (H != 0).sum(0)
    idf = np.log((n - df + 0.5) / (df + 0.5))
    tf = _okapi_ref(H)
    for norm in ("l2", "l1", None):
        ok = OkapiTransformer(compat=False, norm=norm).fit(H)
        np.testing.assert_allclose(ok.idf_, idf, rtol=1e-15)
        want = tf * idf[None, :]
        if norm == "l2":
            want = want / np.sqrt((want ** 2).sum(1, keepdims=True))
        elif norm == "l1":
            want = want / np.abs(want).sum(1, keepdims=True)
        got = ok.transform(H)
        assert sp.issparse(got)
        np.testing.assert_allclose(got.toarray(), want, rtol=1e-14, atol=1e-300)
        dev = ops.require_cuda()
        Hd = torch.from_numpy(H).to(dev)
        got_d = ok.transform(Hd)
        assert got_d.is_cuda and torch.equal(Hd, torch.from_numpy(H).to(dev))     # copy=True
        np.testing.assert_allclose(got_d.cpu().numpy(), want, rtol=1e-14, atol=1e-300)
        got_f = ok.transform(Hd.to(torch.float32)).cpu().numpy()
        np.testing.assert_allclose(got_f, want, rtol=2e-6, atol=1e-30)
    # use_idf=False keeps only the normalisation; the default (compat) ignores both, like the reference
    ok = OkapiTransformer(compat=False, use_idf=False).fit(H)
    np.testing.assert_allclose(ok.transform(H).toarray(), tf / np.sqrt((tf ** 2).sum(1, keepdims=True)), rtol=1e-14)
    np.testing.assert_allclose(OkapiTransformer().fit(H).transform(H).toarray(), tf, rtol=1e-15)


@pytest.mark.parametrize("metric_ip", [True, False])
def test_search_with_k_beyond_the_fused_limit(metric_ip):
    """Faiss's IndexFlat.search takes any k; k > 128 goes through exact pair scores + masked selection passes."""
    from image_search_engine_b200 import faiss_compat
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(11)
    db = unit_rows(rng, 3000, 48)
    db[100] = db[5]                                      # exact tie: lower id first
    q = db[rng.integers(0, 3000, 30)] + 0.05 * rng.standard_normal((30, 48)).astype(np.float32)
    q[0] = db[5]
    idx = faiss_compat.IndexFlatIP(48) if metric_ip else faiss_compat.IndexFlatL2(48)
    idx.add(db)
    for k in (129, 300):
        D, I = idx.search(q, k)
        assert D.shape == (30, k) and I.dtype == np.int64
        Dr, Ir = fs.knn(q, db, k, fs.METRIC_INNER_PRODUCT if metric_ip else fs.METRIC_L2)
        assert_topk_parity(I, Ir, q, db, metric_ip, max_mismatch_frac=0.5)
        np.testing.assert_allclose(D, Dr, rtol=1e-4, atol=1e-5)
        assert list(I[0, :2]) == [5, 100]
        assert all(len(set(row)) == k for row in I)      # no id returned twice across the passes
    D, I = idx.search(q[:3], 3500)                       # k > ntotal > 128: padded like Faiss
    assert (I[:, 3000:] == -1).all() and (I[:, :3000] >= 0).all()
    assert sorted(I[0, :3000]) == list(range(3000))
    pad = -np.finfo(np.float32).max if metric_ip else np.finfo(np.float32).max
    assert (D[:, 3000:] == pad).all()


def test_concurrent_pipelined_transforms_share_one_bovw():
    """Two threads calling transform_csr / histograms_host on ONE BOVW (threaded Flask server, joblib threads):
    the cached staging buffers are guarded, so every call returns its own correct matrix."""
    from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, faiss_compat
    from image_search_engine_b200.bag_of_visual_words import pack_descriptions
    rng = np.random.default_rng(5)
    k, d = 256, 64
    cent = unit_rows(rng, k, d)
    gi = faiss_compat.IndexFlatIP(d)
    gi.add(cent)
    bovw = BOVW(None, n_clusters=k)
    bovw.clusterer = FaissKMeans(k, index=gi)
    batches = []
    for s in range(2):
        descs = [sift_like(rng, int(rng.integers(30, 90)), d) for _ in range(200)]
        batches.append(pack_descriptions(descs, pin=True))
    ok = OkapiTransformer()
    want = [bovw.transform_csr(b, okapi=ok, n_chunks=4).toarray() for b in batches]
    got, errs = {}, []

    def work(t):
        try:
            for it in range(6):
                m = bovw.transform_csr(batches[t], okapi=ok, n_chunks=4, copy=True).toarray()
                out = torch.empty((200, k), dtype=torch.float64, pin_memory=True)
                h = bovw.histograms_host(batches[t], out, okapi=ok, n_chunks=4).copy()
                got[(t, it)] = (m, h)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (t, it), (m, h) in got.items():
        assert np.array_equal(m, want[t]) and np.array_equal(h, want[t]), (t, it)


def test_sharded_index_small_batches_and_empty_shard():
    """world = 1 (no process group): fewer than 20 queries take the same exact kernel as the unsharded index, and a
    shard without rows answers with padded lists instead of raising."""
    from image_search_engine_b200 import faiss_compat
    from image_search_engine_b200._lib import METRIC_L2
    from image_search_engine_b200.parallel import ShardedIndexFlat
    rng = np.random.default_rng(3)
    db = rng.standard_normal((5000, 40)).astype(np.float32)
    q = db[:7] + 1e-4 * rng.standard_normal((7, 40)).astype(np.float32)
    six = ShardedIndexFlat(40, METRIC_L2)
    six.add_local(db)
    flat = faiss_compat.IndexFlatL2(40)
    flat.add(db)
    D, I = six.search(q, 5)
    Dr, Ir = flat.search(q, 5)
    assert np.array_equal(I.cpu().numpy(), Ir) and np.array_equal(D.cpu().numpy(), Dr)   # same kernel, same bits
    empty = ShardedIndexFlat(40, METRIC_L2)
    empty.add_local(np.zeros((0, 40), np.float32))
    D, I = empty.search(q, 5)
    assert (I.cpu().numpy() == -1).all() and (D.cpu().numpy() == np.finfo(np.float32).max).all()
