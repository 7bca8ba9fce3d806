"""Property tests (hypothesis) of the oracle and of the host-side logic: size-independent invariants the GPU parity
tests rely on.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cpu_baseline, faiss_shim as fs

SET = settings(max_examples=25, deadline=None)


@SET
@given(st.integers(1, 60), st.integers(2, 40), st.integers(1, 12), st.booleans(), st.integers(0, 2**31 - 1))
def test_topk_is_permutation_equivariant_and_sorted(nq, nb, k, ip, seed):
    """Permuting the database permutes the returned ids (when the scores are distinct); results are sorted best
    first; k > nb pads with -1."""
    rng = np.random.default_rng(seed)
    d = 7
    x = rng.standard_normal((nq, d)).astype(np.float32)
    y = rng.standard_normal((nb, d)).astype(np.float32)
    metric = fs.METRIC_INNER_PRODUCT if ip else fs.METRIC_L2
    D, I = fs.knn(x, y, k, metric)
    perm = rng.permutation(nb)
    Dp, Ip = fs.knn(x, y[perm], k, metric)
    kk = min(k, nb)
    assert (I[:, kk:] == -1).all()
    srt = -D[:, :kk] if ip else D[:, :kk]
    assert (np.diff(srt, axis=1) >= 0).all()
    distinct = (np.diff(srt, axis=1) > 1e-5).all(axis=1) if kk > 1 else np.ones(nq, bool)
    assert np.array_equal(perm[Ip[distinct, :kk]], I[distinct, :kk])


@SET
@given(st.integers(20, 80), st.integers(30, 200), st.integers(1, 10), st.integers(2, 5), st.integers(0, 2**31 - 1))
def test_merge_of_shards_equals_unsharded_search(nq, nb, k, shards, seed):
    """Row-sharded search + (score, id) merge == search over the whole database (ShardedIndexFlat's contract)."""
    from image_search_engine_b200.parallel import shard_bounds
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((nq, 5)).astype(np.float32)
    y = rng.integers(-3, 4, size=(nb, 5)).astype(np.float32)      # small integers: plenty of exact ties
    D, I = fs.knn(x, y, k, fs.METRIC_INNER_PRODUCT)
    b = shard_bounds(nb, shards)
    parts = []
    for r in range(shards):
        if b[r + 1] > b[r]:
            Dr, Ir = fs.knn(x, y[b[r]:b[r + 1]], k, fs.METRIC_INNER_PRODUCT)
            parts.append((Dr, np.where(Ir >= 0, Ir + b[r], -1)))
    Dm = np.concatenate([p[0] for p in parts], axis=1)
    Im = np.concatenate([p[1] for p in parts], axis=1)
    for q in range(nq):
        valid = Im[q] >= 0
        order = np.lexsort((Im[q][valid], -Dm[q][valid]))[:k]
        assert np.array_equal(Im[q][valid][order], I[q][I[q] >= 0])


@SET
@given(st.lists(st.integers(0, 40), min_size=1, max_size=12), st.integers(2, 64), st.integers(0, 2**31 - 1))
def test_histogram_rows_sum_to_image_sizes(sizes, k, seed):
    """The reference's per-image np.histogram loop (numpy-compat binning, quirk Q1) conserves every descriptor."""
    rng = np.random.default_rng(seed)
    d = 4
    cent = rng.standard_normal((k, d)).astype(np.float32)
    fs.normalize_L2(cent)
    descs = [rng.standard_normal((n, d)).astype(np.float32) for n in sizes if n > 0]
    if not descs:
        return
    H = cpu_baseline.visual_word_histograms(cpu_baseline.codebook_index(cent), descs, k)
    assert np.array_equal(H.sum(1), np.array([len(x) for x in descs], dtype=np.float64))
    T = cpu_baseline.okapi_transform(H)
    assert T.shape == H.shape and ((T.toarray() > 0) == (H > 0)).all() and (T.toarray() < 1).all()


@SET
@given(st.integers(1, 400), st.integers(0, 2**32 - 1))
def test_rand_perm_prefix_is_a_prefix_of_the_permutation(n, seed):
    from image_search_engine_b200 import ops
    full = fs.rand_perm(n, seed)
    assert np.array_equal(np.sort(full), np.arange(n))
    m = max(1, n // 3)
    assert np.array_equal(ops.rand_perm_prefix(n, seed, m), full[:m])
    assert np.array_equal(ops.rand_perm_prefix(n, seed, n), full)


@SET
@given(st.integers(4, 300), st.integers(1, 6), st.integers(0, 2**31 - 1))
def test_split_plan_conserves_points_and_fills_every_cluster(k, n_empty, seed):
    from image_search_engine_b200 import ops
    rng = np.random.default_rng(seed)
    n_empty = min(n_empty, k - 2)
    h = rng.integers(4, 50, k).astype(np.float32)
    h[rng.choice(k, n_empty, replace=False)] = 0
    n = int(h.sum())
    pairs, h2 = ops.split_plan(h, n)
    assert pairs.shape == (n_empty, 2) and (h2 > 0).all() and h2.sum() == pytest.approx(h.sum())
    assert set(pairs[:, 0].tolist()) == set(np.nonzero(h == 0)[0].tolist())


@SET
@given(st.integers(0, 200), st.integers(1, 17))
def test_chunkit_partitions_in_order(n, num):
    from image_search_engine_b200 import chunkIt
    seq = list(range(n))
    pieces = chunkIt(seq, num)
    assert [x for p in pieces for x in p] == seq


@settings(max_examples=200, deadline=None)
@given(st.integers(2, 70000), st.integers(0, 2**31 - 2), st.integers(1, 2**20), st.integers(0, 2**31 - 1))
def test_division_free_bin_estimate_matches_numpy_histogram(k, lo, span, seed):
    """The histogram kernels bin ids with (x - first) * (k / denom) + numpy's own edge corrections instead of
    numpy's ((x - first) / denom) * k (bovw.cu: NumpyBins).  Restated here in float64 NumPy and compared with
    np.histogram over random id ranges, codebook sizes and ids (edges, duplicates, both ends included)."""
    rng = np.random.default_rng(seed)
    hi = min(lo + span, 2**31 - 1)
    ids = rng.integers(lo, hi + 1, 300)
    ids[:2] = [lo, hi]
    ids = np.concatenate([ids, ids[:20]])
    want = np.histogram(ids, bins=k)[0]
    mn, mx = int(ids.min()), int(ids.max())
    first, last = (mn - 0.5, mx + 0.5) if mn == mx else (float(mn), float(mx))
    denom = last - first
    step = denom / k
    kscale = k / denom
    x = ids.astype(np.float64)

    def edge(i):
        return np.where(i == k, last, i.astype(np.float64) * step + first)

    idx = ((x - first) * kscale).astype(np.int64)
    idx[idx == k] -= 1
    idx[x < edge(idx)] -= 1
    inc = (x >= edge(idx + 1)) & (idx != k - 1)
    idx[inc] += 1
    assert idx.min() >= 0 and idx.max() < k
    assert np.array_equal(np.bincount(idx, minlength=k), want)
