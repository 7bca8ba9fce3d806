"""Shared helpers for the parity tests: seeded synthetic inputs and the near-tie rule."""
import numpy as np

EPS32 = float(np.finfo(np.float32).eps)


def orb_like(rng, n, d=32):
    return rng.integers(0, 256, size=(n, d), dtype=np.uint8)


def sift_like(rng, n, d=128):
    g = np.abs(rng.standard_normal((n, d))) ** 2
    g *= 512.0 / np.linalg.norm(g, axis=1, keepdims=True)
    return np.minimum(np.rint(g), 255).astype(np.float32)


def unit_rows(rng, n, d, relu=False):
    x = rng.standard_normal((n, d)).astype(np.float32)
    if relu:
        x = np.maximum(x, 0)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def scores64(x, y, metric_ip):
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    ip = x @ y.T
    if metric_ip:
        return ip
    return (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2 * ip


def near_tie_tol(x, y, factor=16.0):
    """tau = factor * eps_f32 * |x| * max|y|: FP32-level slack on a score, per row."""
    xn = np.linalg.norm(x.astype(np.float64), axis=1)
    yn = np.linalg.norm(y.astype(np.float64), axis=1).max()
    return factor * EPS32 * np.maximum(xn * yn, 1e-30)


def assert_topk_parity(ids, ref_ids, x, y, metric_ip, *, max_mismatch_frac=0.01, factor=16.0):
    """ids must equal ref_ids except where the FP64 scores of the two picks differ by <= tau."""
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    assert ids.shape == ref_ids.shape
    diff = ids != ref_ids
    if not diff.any():
        return 0
    rows = np.nonzero(diff.any(axis=1))[0]
    s = scores64(x[rows], y, metric_ip)
    tol = near_tie_tol(x[rows], y, factor)
    if not metric_ip:
        tol = 2 * tol + factor * EPS32 * ((x[rows].astype(np.float64) ** 2).sum(1) + (y.astype(np.float64) ** 2).sum(1).max())
    for j, r in enumerate(rows):
        cols = np.nonzero(diff[r])[0]
        a, b = ids[r, cols], ref_ids[r, cols]
        assert (a >= 0).all() and (b >= 0).all(), f"row {r}: padding mismatch"
        gap = np.abs(s[j, a] - s[j, b])
        assert (gap <= tol[j]).all(), f"row {r}: ids {a} vs {b} differ beyond near-tie tol (gap {gap.max()}, tol {tol[j]})"
    frac = rows.size / ids.shape[0]
    assert frac <= max_mismatch_frac, f"{frac:.4%} rows differ (near ties only, but too many)"
    return rows.size
