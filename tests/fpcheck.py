"""Shared fp comparison rule of the GPU parity tests (north_star: distances within 1e-5 relative for fp).

scale = |d| for the metrics whose result is a sum of non-negative terms (L2, L2^2, L1); for COSINE 1.0 (the result is
1 - cos) and for DOT the largest |d| of the oracle's top-k (the result is a signed sum: relative to the magnitude of what
is being summed, as in tests/test_gpu_parity.py::test_all_distances_match_oracle).  Rowids must be identical except where
the ORACLE's own distances tie with the k-th value within the same tolerance (then either row is a correct answer)."""
import numpy as np

from oracle import pyoracle as po


def fp_scale(metric, want_d):
    scale = np.maximum(np.abs(want_d), 1e-30)
    if metric == po.COS:
        scale = np.maximum(scale, 1.0)
    elif metric == po.DOT and len(want_d):
        scale = np.maximum(scale, float(np.abs(want_d).max()))
    return scale


def assert_fp_topk(got_ids, got_d, want_ids, want_d, metric, dist_of_row, ctx=()):
    """dist_of_row(rowid) -> the oracle's distance of that row (only called for rows on which the two sides disagree)"""
    assert len(got_ids) == len(want_ids), ctx
    if not len(want_ids):
        return
    scale = fp_scale(metric, want_d)
    err = np.abs(got_d - want_d)
    assert np.all(err <= 1e-5 * scale + 1e-30), (ctx, float((err / scale).max()))
    if not np.array_equal(got_ids, want_ids):
        kth = float(want_d[-1])
        tol = 1e-5 * float(scale.max() if metric in (po.DOT, po.COS) else max(abs(kth), 1e-30))
        diff = set(got_ids.tolist()) ^ set(want_ids.tolist())
        for r in diff:
            assert abs(float(dist_of_row(int(r))) - kth) <= 2 * tol, (ctx, "row", r, "is not a near tie of the k-th distance")
        # the common rows appear in the same relative order unless they tie too
        common = [r for r in got_ids.tolist() if r not in diff]
        wcommon = [r for r in want_ids.tolist() if r not in diff]
        if common != wcommon:
            dmap = dict(zip(want_ids.tolist(), want_d.tolist()))
            for a, b in zip(common, wcommon):
                assert a == b or abs(dmap[a] - dmap[b]) <= 2 * tol, (ctx, "order", a, b)
