"""Parity at the BENCHMARK'S OWN SIZES (BASELINE.json configs[1], [2], [4]-shard) against the CPU oracle, plus the
adversarial inputs for the coarse-pass error bound.

The verified search path is size dependent (column splits, seed pre-pass over a strided sample, soft lock-step,
two row tiles per CTA, fallback re-runs), so small-shape parity does not cover it: here the CUDA path runs the full
configuration and a sample of its rows is re-computed by oracle/faiss_shim.knn on the host (the full database, a
subset of the rows -- a flat search is independent per row).  Tolerances are the parity contract's: ids equal
except FP32 near ties (tests/_util.assert_topk_parity), distances <= 1e-4 relative."""
import numpy as np
import pytest
import torch

from tests._util import EPS32, assert_topk_parity

pytestmark = pytest.mark.gpu


def _sift_like_device(gen, n, d, dev):
    x = torch.randn((n, d), generator=gen, device=dev).square_()
    return torch.minimum((x * (512.0 / x.norm(dim=1, keepdim=True))).round_(), torch.tensor(255.0, device=dev))


def test_c2_full_size_assign_against_oracle():
    """C2: 1 M SIFT-like descriptors x 4096 centroids; 2 000 sampled descriptors re-assigned by the oracle against the
    full codebook; histogram row sums; idempotence."""
    from image_search_engine_b200 import faiss_compat, ops
    from oracle import faiss_shim as fs
    dev = ops.require_cuda()
    gen = torch.Generator(device=dev)
    gen.manual_seed(2)
    n, d, k, per = 1_000_000, 128, 4096, 100
    x = _sift_like_device(gen, n, d, dev)
    cent = x[torch.randperm(n, generator=gen, device=dev)[:k]].clone()
    ops.normalize_l2_(cent)
    idx = faiss_compat.IndexFlatIP(d)
    idx.add(cent)
    dis, words = idx.search(x, 1)
    stats = ops.search_stats()
    rows = torch.randperm(n, generator=gen, device=dev)[:2000]
    xs, cs = x[rows].cpu().numpy(), cent.cpu().numpy()
    Dr, Ir = fs.knn(xs, cs, 1, fs.METRIC_INNER_PRODUCT)
    nm = assert_topk_parity(words[rows].cpu().numpy(), Ir, xs, cs, True, max_mismatch_frac=0.002)
    np.testing.assert_allclose(dis[rows].cpu().numpy(), Dr, rtol=1e-4)
    print(f"C2 full size: mode={stats['mode']} fallback_rows={stats['fallback_rows']} near-tie rows in sample={nm}/2000")
    off = torch.arange(0, n + 1, per, device=dev, dtype=torch.int64)
    H = ops.bovw_histogram(words.reshape(-1), off, k)
    assert float(H.sum()) == n and bool((H.sum(1) == per).all())
    # the histogram of the sampled images equals np.histogram on the oracle-checked words
    wh = words.reshape(-1).cpu().numpy()
    for img in (0, 1234, 9999):
        ref, _ = np.histogram(wh[img * per:(img + 1) * per], bins=k)
        assert np.array_equal(H[img].cpu().numpy(), ref.astype(np.float64))


def _c3_like(dev, nb, d, nq, seed, relu=True, noise=0.05):
    from image_search_engine_b200 import ops
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    db = torch.empty((nb, d), dtype=torch.float32, device=dev)
    for i in range(0, nb, 100_000):
        blk = db[i:i + 100_000]
        blk.normal_(generator=g)
        if relu:
            blk.clamp_(min=0)
    ops.normalize_l2_(db)
    pick = torch.randint(0, nb, (nq,), generator=g, device=dev)
    q = db[pick] + noise * torch.randn((nq, d), generator=g, device=dev)
    ops.normalize_l2_(q)
    return db, q, pick


def test_c3_full_size_search_against_oracle():
    """C3: 1 M x 2048 inner-product index, 10 k queries, top-10 on the verified coarse path; 200 sampled queries are
    searched again by the oracle over the FULL database on the host."""
    from image_search_engine_b200 import faiss_compat, ops
    from oracle import faiss_shim as fs
    dev = ops.require_cuda()
    db, q, pick = _c3_like(dev, 1_000_000, 2048, 10_000, 3)
    idx = faiss_compat.IndexFlatIP(2048)
    idx.add(db)
    D, I = idx.search(q, 10)
    stats = ops.search_stats()
    assert stats["mode"] == "verified"
    assert float((I[:, 0] == pick).float().mean()) > 0.999
    g = torch.Generator(device=dev)
    g.manual_seed(33)
    rows = torch.randperm(10_000, generator=g, device=dev)[:200]
    qs, dbh = q[rows].cpu().numpy(), db.cpu().numpy()
    del idx
    Dr, Ir = fs.knn(qs, dbh, 10, fs.METRIC_INNER_PRODUCT, db_block=65536)
    nm = assert_topk_parity(I[rows].cpu().numpy(), Ir, qs, dbh, True, max_mismatch_frac=0.05)
    np.testing.assert_allclose(D[rows].cpu().numpy(), Dr, rtol=1e-4, atol=1e-6)
    print(f"C3 full size: fallback_rows={stats['fallback_rows']} of {stats['rows']}, near-tie rows in sample={nm}/200")


def test_c5_shard_top100_collect_mode_against_oracle():
    """One C5 shard: 1.25 M x 512, 10 k queries, top-100 (collect mode); 100 sampled queries against the oracle."""
    from image_search_engine_b200 import faiss_compat, ops
    from oracle import faiss_shim as fs
    dev = ops.require_cuda()
    db, q, pick = _c3_like(dev, 1_250_000, 512, 10_000, 5, relu=False)
    idx = faiss_compat.IndexFlatIP(512)
    idx.add(db)
    D, I = idx.search(q, 100)
    stats = ops.search_stats()
    assert stats["mode"] == "verified-collect"
    assert float((I[:, 0] == pick).float().mean()) > 0.999
    assert bool((D[:, :-1] >= D[:, 1:]).all()) and bool((I >= 0).all())
    g = torch.Generator(device=dev)
    g.manual_seed(55)
    rows = torch.randperm(10_000, generator=g, device=dev)[:100]
    qs, dbh = q[rows].cpu().numpy(), db.cpu().numpy()
    Dr, Ir = fs.knn(qs, dbh, 100, fs.METRIC_INNER_PRODUCT, db_block=65536)
    nm = assert_topk_parity(I[rows].cpu().numpy(), Ir, qs, dbh, True, max_mismatch_frac=0.2)
    np.testing.assert_allclose(D[rows].cpu().numpy(), Dr, rtol=1e-4, atol=1e-6)
    print(f"C5 shard: fallback_rows={stats['fallback_rows']} of {stats['rows']}, near-tie rows in sample={nm}/100")


def _adversarial(rng, m, n, d, lo_exp=-16, hi_exp=-9):
    """All-positive operands whose elements sit at the top of the scaled FP16 range (every product has the same
    sign: the truncating tensor-core accumulation errs in one direction and the partial sums are as large as they
    get), and a database of near-duplicates whose exact scores differ by about the coarse error bound."""
    a = rng.uniform(0.90, 1.0, size=(m, d)).astype(np.float32)
    base = rng.uniform(0.90, 1.0, size=(8, d)).astype(np.float32)
    # relative perturbations from 2^-16 to 2^-9 around the 2^-11 FP16 rounding step of the planes
    scale = np.exp2(rng.uniform(lo_exp, hi_exp, size=(n, 1))).astype(np.float32)
    b = base[rng.integers(0, 8, n)] * (1.0 + scale * rng.standard_normal((n, d)).astype(np.float32))
    b[1::97] = b[0::97][: len(b[1::97])]                 # exact duplicates: lower id must win
    return a, np.ascontiguousarray(b, dtype=np.float32)


@pytest.mark.parametrize("spread", ["at_the_bound", "above_the_bound"])
@pytest.mark.parametrize("metric_ip", [True, False])
def test_coarse_bound_adversarial_topk(metric_ip, spread):
    """d = 2048: every row the proof accepts must equal the FP32-grade split path and the FP64 ground truth (up to
    FP32 near ties); the bound is allowed to be pessimistic (fallback rows), never optimistic."""
    from image_search_engine_b200 import ops
    from image_search_engine_b200._lib import METRIC_IP, METRIC_L2
    rng = np.random.default_rng(2048)
    m, n, d, k = 256, 6000, 2048, 10
    # "at_the_bound": near-duplicate columns (perturbations of 2^-16 .. 2^-9 around the 2^-11 rounding step): nothing can
    # be proven, every row must fall back
    a, b = _adversarial(rng, m, n, d)
    if spread == "above_the_bound":
        # a head of 40 columns whose scores step down by 1 % of |a||b| (ten times the bound), the rest far below: the
        # proof accepts these rows -- and they are exactly the rows an optimistic bound would get wrong
        gain = np.full((n, 1), 0.5, np.float32)
        head = rng.choice(n, 40, replace=False)
        gain[head, 0] = 1.0 - 0.01 * np.arange(40, dtype=np.float32)
        gain[gain[:, 0] == 0.5, 0] *= rng.uniform(0.5, 1.0, size=n - 40).astype(np.float32)
        b = np.ascontiguousarray(b * gain, dtype=np.float32)
    dev = ops.require_cuda()
    ad, bd = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    a_op, b_op = ops.prepare_operand(ad), ops.attach_sample(ops.prepare_operand(bd))
    metric = METRIC_IP if metric_ip else METRIC_L2
    Dv, Iv = ops.search_topk(ad, a_op, bd, b_op, metric, k, precision="verified")
    sv = ops.search_stats()
    Ds, Is = ops.search_topk(ad, a_op, bd, b_op, metric, k, precision="split")
    assert sv["mode"] == "verified"
    print(f"adversarial top-{k} ({'IP' if metric_ip else 'L2'}, {spread}): fallback_rows={sv['fallback_rows']} of {m}")
    if spread == "above_the_bound":
        assert sv["fallback_rows"] < m, "no row was accepted by the proof: the accepted-row check below is vacuous"
    # ground truth in float64
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    s = a64 @ b64.T
    if not metric_ip:
        s = (a64 ** 2).sum(1)[:, None] + (b64 ** 2).sum(1)[None, :] - 2 * s
    key = -s if metric_ip else s
    order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), key), axis=1)[:, :k]
    assert_topk_parity(Iv.cpu().numpy(), order, a, b, metric_ip, max_mismatch_frac=1.0)
    assert_topk_parity(Iv.cpu().numpy(), Is.cpu().numpy(), a, b, metric_ip, max_mismatch_frac=1.0)
    # completeness, stated directly: nothing outside the returned list may beat the returned k-th by more than FP32 noise
    got = Iv.cpu().numpy()
    kth = np.take_along_axis(s, got[:, -1:], 1)[:, 0]
    # the same FP32 near-tie slack as assert_topk_parity (tau = 16 eps |a| max|b|; L2 adds the norms' rounding)
    tol = 16.0 * EPS32 * np.linalg.norm(a64, axis=1) * np.linalg.norm(b64, axis=1).max()
    if not metric_ip:
        tol = 2 * tol + 16.0 * EPS32 * ((a64 ** 2).sum(1) + (b64 ** 2).sum(1).max())
    mask = np.ones_like(s, dtype=bool)
    np.put_along_axis(mask, got, False, 1)
    outside_best = np.where(mask, s, -np.inf if metric_ip else np.inf)
    outside_best = outside_best.max(1) if metric_ip else outside_best.min(1)
    bad = (outside_best > kth + tol) if metric_ip else (outside_best < kth - tol)
    assert not bad.any(), f"{bad.sum()} rows miss a true neighbour"


def test_coarse_bound_adversarial_top1_verification():
    """The verified TOP-1 path (coarse winner + exact runner-up, rows flagged when the gap is inside the bound) on the
    same adversarial data: enough rows that the path is actually taken (>= 4 row tiles per SM), FP64 check on a sample."""
    from image_search_engine_b200 import _lib, ops
    from image_search_engine_b200._lib import METRIC_IP
    dev = ops.require_cuda()
    sms = _lib.load().ise_ctx_sm_count(_lib.ctx(dev.index))
    rng = np.random.default_rng(4096)
    m, n, d = 4 * 128 * sms, 1024, 2048
    a_small, b = _adversarial(rng, 4096, n, d)
    ad = torch.from_numpy(a_small).to(dev).repeat((m + 4095) // 4096, 1)[:m].contiguous()
    ad *= 1.0 + 2e-4 * torch.randn((m, 1), device=dev)       # rows differ, structure stays
    bd = torch.from_numpy(b).to(dev)
    a_op, b_op = ops.prepare_operand(ad), ops.prepare_operand(bd)
    Dv, Iv = ops.search_topk(ad, a_op, bd, b_op, METRIC_IP, 1, precision="verified")
    sv = ops.search_stats()
    assert sv["mode"] == "verified" and sv["rows"] == m
    print(f"adversarial top-1: fallback_rows={sv['fallback_rows']} of {m}")
    Ds, Is = ops.search_topk(ad, a_op, bd, b_op, METRIC_IP, 1, precision="split")
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    rows = torch.randperm(m, generator=g, device=dev)[:3000]
    a_s = ad[rows].cpu().numpy()
    s = a_s.astype(np.float64) @ b.astype(np.float64).T
    truth = s.argmax(1)[:, None]
    assert_topk_parity(Iv[rows].cpu().numpy(), truth, a_s, b, True, max_mismatch_frac=1.0)
    assert_topk_parity(Is[rows].cpu().numpy(), truth, a_s, b, True, max_mismatch_frac=1.0)
    same = (Iv == Is).float().mean().item()
    assert same > 0.9, same
