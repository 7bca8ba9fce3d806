"""CPU: the C-ABI library loads and exports every symbol include/ise.h declares, the host-side
helpers match the oracle, and the product path fails loudly without a GPU."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import ctypes
    from image_search_engine_b200 import _lib
    header = (ROOT / "include" / "ise.h").read_text()
    declared = set(re.findall(r"\b(ise_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libise.so"
    assert _lib.load().ise_version() == 100


def test_rand_perm_prefix_matches_oracle():
    from image_search_engine_b200 import ops
    from oracle import faiss_shim as fs
    for n, seed, m in [(10, 1, 10), (1, 5, 1), (1000, 43, 1000), (500000, 42 + 1, 512), (131072 * 4, 42, 131072)]:
        assert np.array_equal(ops.rand_perm_prefix(n, seed, m), fs.rand_perm(n, seed, prefix=m))
    assert np.array_equal(ops.rand_perm_prefix(50, 2**32 + 9, 50), ops.rand_perm_prefix(50, 9, 50))  # (unsigned)seed


def test_split_plan_matches_oracle():
    from image_search_engine_b200 import ops
    from oracle import faiss_shim as fs
    rng = np.random.default_rng(4)
    # the last case has n >> sum(h): tiny donor probabilities, thousands of RNG draws per split (the regime of a
    # large codebook) -- the integer-threshold form of the test in ise_split_plan must take the same decisions
    for k, n_empty, n_over in [(16, 3, 1), (200, 40, 1), (1000, 1, 1), (2000, 60, 20)]:
        h = rng.integers(1, 60, k).astype(np.float32)
        h[rng.choice(k, n_empty, replace=False)] = 0
        n = int(h.sum()) * n_over
        cent = rng.standard_normal((k, 6)).astype(np.float32)
        cent[h == 0] = 0
        c_ref, h_ref = cent.copy(), h.copy()
        ns = fs.split_clusters(6, k, n, h_ref, c_ref)
        pairs, h_new = ops.split_plan(h, n)
        assert pairs.shape == (ns, 2) and ns == n_empty
        assert np.array_equal(h_new, h_ref)
        # replay the plan on the host exactly as the device kernel does
        c = cent.copy()
        up, dn = np.float32(1 + 1 / 1024), np.float32(1 - 1 / 1024)
        for ci, cj in pairs:
            v = c[cj].copy()
            c[ci, 0::2], c[cj, 0::2] = v[0::2] * up, v[0::2] * dn
            c[ci, 1::2], c[cj, 1::2] = v[1::2] * dn, v[1::2] * up
        assert np.array_equal(c, c_ref)


def test_chunkit_and_okapi_fit_host_logic():
    from image_search_engine_b200 import OkapiTransformer, chunkIt
    g = np.load(ROOT / "tests" / "golden" / "bovw_c1mini.npz")
    assert [len(c) for c in chunkIt(list(range(37)), 5)] == list(g["chunk_bounds"])
    assert sum(chunkIt(list(range(10)), 3), []) == list(range(10))
    ok = OkapiTransformer().fit(g["hist"])
    np.testing.assert_allclose(ok.idf_, g["idf"], rtol=1e-12)
    ok.idf_ = np.arange(4.0)
    assert list(ok.idf_) == [0, 1, 2, 3]
    from sklearn.base import clone
    assert clone(OkapiTransformer(b=0.5)).b == 0.5


def test_estimators_are_clonable_and_light():
    from sklearn.base import clone
    from image_search_engine_b200 import BOVW, FaissKMeans
    b = clone(BOVW(describer=None, n_clusters=17))
    assert b.n_clusters == 17 and b.get_params()["hist_mode"] == "numpy_compat"
    b.set_params(n_clusters=33)
    assert b.n_clusters == 33
    km = FaissKMeans()
    assert (km.n_clusters, km.n_init, km.max_iter, km.init_centroids, km.index) == (8, 3, 25, None, None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from image_search_engine_b200 import FaissKMeans, IseError, create_search_index, faiss_compat
    x = np.zeros((64, 8), np.float32)
    with pytest.raises(IseError):
        FaissKMeans(4).fit(x)
    with pytest.raises(IseError):
        create_search_index(x.copy(), "l2")
    with pytest.raises(IseError):
        faiss_compat.normalize_L2(x)


def test_product_code_never_imports_the_oracle():
    pkg = ROOT / "image_search_engine_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert "oracle" not in f.read_text()


def test_shard_bounds():
    from image_search_engine_b200.parallel import shard_bounds
    assert list(shard_bounds(10, 4)) == [0, 3, 6, 8, 10]
    assert list(shard_bounds(8, 8)) == list(range(9))
    assert shard_bounds(10_000_000, 8)[-1] == 10_000_000


def test_query_batcher_host_logic():
    """QueryBatcher's batching / padding / future plumbing with a NumPy stand-in for the index (no GPU): every
    request gets its own top-n, batches are padded to Faiss's BLAS threshold so one code path serves them all,
    a failing search fails its requests and not the server thread."""
    from image_search_engine_b200.engine import QueryBatcher

    class FakeIndex:
        def __init__(self, db):
            self.db, self.calls = db, []

        def search(self, q, k):
            self.calls.append(q.shape[0])
            if np.isnan(q).any():
                raise RuntimeError("bad query")
            d = ((q[:, None, :] - self.db[None]) ** 2).sum(-1)
            I = np.argsort(d, axis=1, kind="stable")[:, :k]
            return np.take_along_axis(d, I, 1).astype(np.float32), I.astype(np.int64)

    rng = np.random.default_rng(0)
    db = rng.standard_normal((50, 8)).astype(np.float32)
    idx = FakeIndex(db)
    paths = [f"p{i}" for i in range(50)]
    with QueryBatcher(idx, paths, max_batch=64, max_wait_ms=150.0) as qb:
        futs = [qb.submit(db[i] + 0.01, 3 + i % 2) for i in range(7)]
        got = [f.result(timeout=30) for f in futs]
        assert sum(qb.batches) == 7
        bad = qb.submit(np.full(8, np.nan, np.float32), 3)
        with pytest.raises(RuntimeError):
            bad.result(timeout=30)
        assert qb.query(db[11], 1)[0][2] == "p11"            # the server thread survived the failed batch
    assert all(c >= 20 for c in idx.calls)                    # padded to distance_compute_blas_threshold
    for i, preds in enumerate(got):
        assert len(preds) == 3 + i % 2 and preds[0][2] == f"p{i}" and preds[0][1] is None
        assert [p[0] for p in preds] == sorted(p[0] for p in preds)


def test_cluster_score_host_logic_matches_reference_fixture():
    """utils.calc_sampled_cluster_score with the oracle's flat-IP search standing in for the GPU quantiser: same
    RandomState(42) stream and sklearn calls as the reference -> the reference's own scores (golden fixture)."""
    import types
    from pathlib import Path
    from image_search_engine_b200 import utils
    from oracle import faiss_shim as fs
    gold_dir = Path(__file__).parent / "golden"
    gold, g = np.load(gold_dir / "cluster_score.npz"), np.load(gold_dir / "bovw_c1mini.npz")
    oidx = fs.IndexFlatIP(32)
    oidx.add(gold["centroids"])
    clusterer = types.SimpleNamespace(transform=lambda X: oidx.search(np.asarray(X, dtype=np.float32), 1)[1])
    off = g["offsets"]
    descs = [g["X"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    est = types.SimpleNamespace(named_steps={"bovw": types.SimpleNamespace(descriptions=descs, clusterer=clusterer)})
    utils.rs = np.random.RandomState(42)
    utils.CLUSTER_EVAL_SAMPLE_SIZE, utils.CLUSTER_EVAL_N_SAMPLES = 2000, 10
    assert utils.calc_sampled_cluster_score(est, None) == pytest.approx(float(gold["score_first_call"]), rel=1e-12)
    assert utils.calc_sampled_cluster_score(est, None) == pytest.approx(float(gold["score_second_call"]), rel=1e-12)


def test_train_bovw_model_keeps_the_reference_signature(monkeypatch):
    """indexer.py:37 calls train_bovw_model(images_paths, describer): config must be optional and, when omitted,
    resolve to the host application's own config.Config() like bag_of_visual_words.py:32,37 does."""
    import inspect
    import sys
    import types
    from image_search_engine_b200 import train_bovw_model
    params = list(inspect.signature(train_bovw_model).parameters.values())
    assert [p.name for p in params[:2]] == ["images_paths", "describer"]
    assert all(p.default is not inspect.Parameter.empty for p in params[2:])
    # no `config` module on sys.path -> a clear error, not a TypeError about a missing argument
    monkeypatch.setitem(sys.modules, "config", None)
    with pytest.raises(RuntimeError, match="config"):
        train_bovw_model(np.zeros((0, 1)), None)
    # an application config module is picked up (the call then proceeds to the pipeline and needs data / a GPU)
    seen = {}

    class Config:
        NUM_CLUSTERS = 3

        def __init__(self):
            seen["made"] = True
    monkeypatch.setitem(sys.modules, "config", types.SimpleNamespace(Config=Config))
    with pytest.raises(Exception):
        train_bovw_model(np.zeros((0, 1)), None)
    assert seen.get("made")


def test_bovw_pickle_does_not_strip_the_live_estimator():
    """BaseEstimator.__getstate__ returns the live __dict__ on Python >= 3.11: pickling must not delete the cached
    pipelines of the estimator being pickled, and must not persist them either."""
    import pickle
    from image_search_engine_b200 import BOVW
    b = BOVW(None, n_clusters=5)
    b.__dict__["_pipe_cache"] = {"key": "x"}
    b.__dict__["_csr_bufs"] = {"cap": 1}
    b._lock()
    blob = pickle.dumps(b)
    assert "_pipe_cache" in b.__dict__ and "_csr_bufs" in b.__dict__ and "_pipe_lock" in b.__dict__
    c = pickle.loads(blob)
    assert c.n_clusters == 5 and not {"_pipe_cache", "_csr_bufs", "_pipe_lock"} & set(c.__dict__)
    with c._lock():          # re-created lazily, re-entrant
        with c._lock():
            pass


def test_okapi_params_and_opt_in_flags():
    from sklearn.base import clone
    from image_search_engine_b200 import OkapiTransformer
    ok = OkapiTransformer()
    assert ok.get_params() == dict(norm="l2", use_idf=True, k1=1, k2=1, b=0.75, compat=True)
    assert ok._norm_code() == 0                        # reference behaviour: norm is declared, never applied
    ok2 = clone(OkapiTransformer(compat=False, norm="l1"))
    assert ok2._norm_code() == 1 and OkapiTransformer(compat=False)._norm_code() == 2
    assert OkapiTransformer(compat=False, norm=None)._norm_code() == 0
    with pytest.raises(ValueError):
        OkapiTransformer(compat=False, norm="max")._norm_code()


def test_split_plan_sparse_stream_index_equals_dense_scan(monkeypatch):
    """Large codebook regime (k = 65 536, C4): the plan walks the cached sparse index of the mt19937(1234) stream; it
    must take exactly the decisions of the draw-by-draw scan (ISE_SPLIT_PLAN_DENSE=1), on a cold index, on a warm
    one, after the background warm-up, and when a later call needs more of the stream than an earlier one."""
    import time
    from image_search_engine_b200 import _lib, ops
    rng = np.random.default_rng(65536)
    k = 65536
    h = rng.poisson(150, k).astype(np.float32) + 2
    h[rng.choice(k, 40, replace=False)] *= 20                     # a few big clusters (still p_max < 2^-8)
    n = 10_000_000

    def plan(n_empty, dense):
        hh = h.copy()
        hh[np.random.default_rng(n_empty).choice(k, n_empty, replace=False)] = 0
        if dense:
            monkeypatch.setenv("ISE_SPLIT_PLAN_DENSE", "1")
        else:
            monkeypatch.delenv("ISE_SPLIT_PLAN_DENSE", raising=False)
        t0 = time.perf_counter()
        pairs, h_new = ops.split_plan(hh, n)
        return pairs, h_new, time.perf_counter() - t0

    p_small, h_small, t_cold = plan(30, dense=False)             # cold index: generates ~2 M draws
    _lib.check(_lib.load().ise_split_plan_warm(1_000_000))        # a finished warm-up must not block a later, larger one
    time.sleep(0.2)
    _lib.check(_lib.load().ise_split_plan_warm(40_000_000))       # background warm-up, returns at once
    p_big, h_big, _ = plan(400, dense=False)                      # needs ~26 M draws: extends / waits for the warm-up
    p_big2, h_big2, t_warm = plan(400, dense=False)               # fully cached now
    d_small, dh_small, _ = plan(30, dense=True)
    d_big, dh_big, t_dense = plan(400, dense=True)
    assert p_small.shape == (30, 2) and p_big.shape == (400, 2)
    assert np.array_equal(p_small, d_small) and np.array_equal(h_small, dh_small)
    assert np.array_equal(p_big, d_big) and np.array_equal(h_big, dh_big)
    assert np.array_equal(p_big, p_big2) and np.array_equal(h_big, h_big2)
    print(f"split plan, k=65536, 400 splits: dense scan {t_dense * 1e3:.1f} ms, cached sparse index {t_warm * 1e3:.2f} ms")
    assert t_warm < t_dense


def test_host_packer_whole_and_chunk_ordered_background_job():
    """ise_pack_rows / ise_pack_begin-wait-end (pure host code): the packed matrix equals np.concatenate
    (bag_of_visual_words.py:128) for ragged lists with empty images; float32 -> uint8 narrowing is accepted exactly when
    every value is an integer in [0, 255]; the background job reports a value that does not fit from whichever chunk
    holds it."""
    import ctypes as C
    from image_search_engine_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    d = 32
    sizes = rng.integers(0, 90, 257)
    sizes[[0, 100, 256]] = 0
    imgs = [rng.integers(0, 256, (int(s), d)).astype(np.float32) for s in sizes]
    want = np.concatenate(imgs)
    ptrs = np.array([a.ctypes.data for a in imgs], dtype=np.uint64)
    offsets = np.zeros(len(imgs) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    p = lambda a: C.c_void_p(a.ctypes.data)

    def whole(dst_dtype, nthreads):
        dst = np.zeros(want.shape, dtype=np.uint8 if dst_dtype == _lib.DTYPE_U8 else np.float32)
        ok = C.c_int(-1)
        _lib.check(lib.ise_pack_rows(p(ptrs), p(offsets), 0, len(imgs), d, _lib.DTYPE_F32, dst_dtype, p(dst), nthreads, C.byref(ok)))
        return dst, ok.value

    def job(dst_dtype, cuts, nthreads):
        dst = np.zeros(want.shape, dtype=np.uint8 if dst_dtype == _lib.DTYPE_U8 else np.float32)
        cuts = np.asarray(cuts, dtype=np.int64)
        h = C.c_void_p()
        _lib.check(lib.ise_pack_begin(p(ptrs), p(offsets), p(cuts), len(cuts) - 1, d, _lib.DTYPE_F32, dst_dtype, p(dst),
                                      nthreads, C.byref(h)))
        oks = []
        for c in range(len(cuts) - 1):
            ok = C.c_int(-1)
            _lib.check(lib.ise_pack_wait(h, c, C.byref(ok)))
            if ok.value:                      # chunk c is in place as soon as its wait returns
                r0, r1 = offsets[cuts[c]], offsets[cuts[c + 1]]
                assert np.array_equal(dst[r0:r1].astype(np.float32), want[r0:r1])
            oks.append(ok.value)
        _lib.check(lib.ise_pack_end(h))
        return dst, oks

    for nt in (1, 3, 8):
        for dt in (_lib.DTYPE_F32, _lib.DTYPE_U8):
            dst, ok = whole(dt, nt)
            assert ok == 1 and np.array_equal(dst.astype(np.float32), want)
            dst, oks = job(dt, [0, 1, 1, 60, 200, 257], nt)
            assert all(oks) and np.array_equal(dst.astype(np.float32), want)
    # one value of image 230 is not a byte: uint8 refuses (whole call and the job), float32 is unaffected
    imgs[230][1, 5] = 17.5
    want = np.concatenate(imgs)
    assert whole(_lib.DTYPE_U8, 4)[1] == 0
    assert whole(_lib.DTYPE_F32, 4)[1] == 1
    _, oks = job(_lib.DTYPE_U8, [0, 64, 128, 192, 257], 4)
    assert oks[-1] == 0
    dst, oks = job(_lib.DTYPE_F32, [0, 64, 128, 192, 257], 4)
    assert all(oks) and np.array_equal(dst, want)
    for bad in (-1.0, 256.0, np.nan):
        imgs[230][1, 5] = bad
        assert whole(_lib.DTYPE_U8, 2)[1] == 0
    # two-ended use: a chunk the caller claims is skipped by the workers (its rows stay untouched), every other chunk is
    # packed; polling reports completion without blocking
    imgs[230][1, 5] = 3.0
    want = np.concatenate(imgs)
    cuts = np.asarray([0, 64, 128, 192, 257], dtype=np.int64)
    dst = np.full(want.shape, 255, dtype=np.uint8)
    h = C.c_void_p()
    _lib.check(lib.ise_pack_begin(p(ptrs), p(offsets), p(cuts), 4, d, _lib.DTYPE_F32, _lib.DTYPE_U8, p(dst), 2, C.byref(h)))
    got = C.c_int(-1)
    _lib.check(lib.ise_pack_claim(h, 3, C.byref(got)))
    claimed = bool(got.value)
    for c in range(4):
        ok = C.c_int(-1)
        _lib.check(lib.ise_pack_wait(h, c, C.byref(ok)))
        assert ok.value == 1
        done, ok2 = C.c_int(-1), C.c_int(-1)
        _lib.check(lib.ise_pack_poll(h, c, C.byref(done), C.byref(ok2)))
        assert ok2.value == 1 and bool(done.value) == (not (claimed and c == 3))
    _lib.check(lib.ise_pack_claim(h, 0, C.byref(got)))
    assert got.value == 0                                    # the workers had it: too late to claim
    _lib.check(lib.ise_pack_end(h))
    r3 = offsets[cuts[3]]
    assert np.array_equal(dst[:r3].astype(np.float32), want[:r3])
    if claimed:
        assert (dst[r3:] == 255).all()
    else:
        assert np.array_equal(dst[r3:].astype(np.float32), want[r3:])
