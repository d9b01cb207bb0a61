"""Comparators shared by the parity tests (rules: SURVEY 8c, DESIGN.md "Parity")."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REL_TOL = 1e-4   # north_star: "within a stated FP32 tolerance, e.g. rel 1e-4"
ABS_FLOOR = 1e-6


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close(a, b, rel=REL_TOL, floor=ABS_FLOOR, what=""):
    """|a-b| <= rel*max(|a|,|b|) + floor, elementwise (SURVEY 8c rule 4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = rel * np.maximum(np.abs(a), np.abs(b)) + floor
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), a.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside rel {rel}; worst at {i}: {a[i]} vs {b[i]}")


def frac_close(a, b, rel=REL_TOL, floor=ABS_FLOOR):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)) + floor).mean())


def assert_knn_equal_mod_ties(idx, ref_idx, ref_dist, what=""):
    """Exact-tie groups compare as sets (SURVEY 8c rule 2): torch.topk orders exact ties
    arbitrarily, ours is lowest-index-first.  ref_dist: (B,N,N) reference distances.
    Column 0 of the (k+1) selection is dropped positionally by both sides, so the
    comparison is on the multiset of *distances* per rank plus index equality off-tie."""
    idx = np.asarray(idx, dtype=np.int64)
    ref_idx = np.asarray(ref_idx, dtype=np.int64)
    assert idx.shape == ref_idx.shape, what
    B, N, k = idx.shape
    for b in range(B):
        for n in range(N):
            if (idx[b, n] == ref_idx[b, n]).all():
                continue
            d_mine = ref_dist[b, n, idx[b, n]]
            d_ref = ref_dist[b, n, ref_idx[b, n]]
            # same distances rank by rank, bit for bit
            assert (d_mine.view(np.uint32) == d_ref.view(np.uint32)).all() or np.array_equal(d_mine, d_ref), \
                f"{what}: row ({b},{n}) differs beyond exact ties: {idx[b, n]} vs {ref_idx[b, n]}"


def knn_feat_mismatch(idx, ref_idx, ref_dist, D, q=None):
    """Feature-space rule (SURVEY 8c rule 3): an index mismatch is accepted only if the swapped
    candidates' reference distances differ by < 8*2^-24*sqrt(D)*(|q_i|+|q_j|).
    Returns (rows_any_change, rows_set_change, violations)."""
    idx = np.asarray(idx, dtype=np.int64)
    ref_idx = np.asarray(ref_idx, dtype=np.int64)
    B, N, k = idx.shape
    rows_any = rows_set = viol = 0
    for b in range(B):
        qq = q[b] if q is not None else None
        for n in range(N):
            if (idx[b, n] == ref_idx[b, n]).all():
                continue
            rows_any += 1
            if set(idx[b, n].tolist()) != set(ref_idx[b, n].tolist()):
                rows_set += 1
            d_mine = ref_dist[b, n, idx[b, n]].astype(np.float64)
            d_ref = ref_dist[b, n, ref_idx[b, n]].astype(np.float64)
            if qq is not None:
                scale = np.abs(qq[n]) + np.maximum(np.abs(qq[idx[b, n]]), np.abs(qq[ref_idx[b, n]]))
            else:
                scale = np.abs(ref_dist[b, n, n]) + 1.0
            bound = 8 * 2.0 ** -24 * np.sqrt(D) * scale
            if (np.abs(d_mine - d_ref) > bound).any():
                viol += 1
    return rows_any, rows_set, viol


def assert_grad_close(a, b, rel=REL_TOL, scale_floor=2e-5, what=""):
    """Gradient comparator: |a-b| <= rel*max(|a|,|b|) + scale_floor*max|b|.  Gradients are long signed sums
    (over B*N*S rows) whose small entries are cancellation residues of O(max|b|) terms, so the absolute floor is
    tied to the tensor's scale instead of the fixed 1e-6 used for activations."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isfinite(a).all(), f"{what}: non-finite gradient"
    err = np.abs(a - b)
    tol = rel * np.maximum(np.abs(a), np.abs(b)) + scale_floor * max(float(np.abs(b).max()), 1e-30)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), a.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside tolerance; worst at {i}: {a[i]} vs {b[i]} "
                             f"(scale {np.abs(b).max():.3e})")
