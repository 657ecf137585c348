"""Batched queries (tcgen05 tensor-core path) vs the oracle and vs the per-query CUDA-core path.  -m gpu."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
METRICS = [po.L2, po.L2SQ, po.COS, po.DOT]


def _index(vtype, x, rowids=None):
    import sqlite_vector_b200 as vs
    ix = vs.Index(vtype, x.shape[1], x.shape[0])
    ix.append_dense(x, rowids)
    ix.finalize()
    return ix


@pytest.mark.parametrize("vtype", [po.I8, po.U8])
@pytest.mark.parametrize("metric", METRICS)
def test_batch_int_bit_exact(oracle, vtype, metric):
    import sqlite_vector_b200 as vs
    eng = vs.load_engine()
    rng = np.random.Generator(np.random.PCG64(500 + 10 * vtype + metric))
    for (n, dim, nq, k) in [(20000, 128, 16, 20), (50000, 384, 100, 20), (30000, 200, 300, 33)]:
        x = po.convert(rng.standard_normal((n, dim), dtype=np.float32), vtype)
        q = po.convert(rng.standard_normal((nq, dim), dtype=np.float32), vtype)
        rowids = np.arange(n, dtype=np.int64) * 2 + 5
        ix = _index(vtype, x, rowids)
        b0 = ix.stat("batches")
        res = ix.scan_topk(metric, q, k)
        assert ix.stat("batches") == b0 + 1, "the tensor-core batch path did not run"
        for b in list(range(0, nq, max(1, nq // 7)))[:8] + [nq - 1]:
            want_ids, want_d = oracle.scan_dense(metric, vtype, q[b], x, rowids, k)
            assert np.array_equal(res[b][0], want_ids), (vtype, metric, n, dim, nq, b)
            assert np.array_equal(res[b][1], want_d)
        eng.set_option("no_batch", 1)
        try:
            loop = ix.scan_topk(metric, q, k)
        finally:
            eng.set_option("no_batch", 0)
        for b in range(nq):
            assert np.array_equal(res[b][0], loop[b][0]) and np.array_equal(res[b][1], loop[b][1]), (vtype, metric, b)
        ix.close()


@pytest.mark.parametrize("vtype", [po.BF16, po.F16])
@pytest.mark.parametrize("metric", METRICS)
def test_batch_fp_matches_single_query_path(oracle, vtype, metric):
    import sqlite_vector_b200 as vs
    eng = vs.load_engine()
    rng = np.random.Generator(np.random.PCG64(700 + 10 * vtype + metric))
    for (n, dim, nq, k) in [(20000, 128, 32, 20), (40000, 768, 130, 20)]:
        x = po.convert(rng.standard_normal((n, dim), dtype=np.float32), vtype)
        q = po.convert(rng.standard_normal((nq, dim), dtype=np.float32), vtype)
        ix = _index(vtype, x)
        b0 = ix.stat("batches")
        res = ix.scan_topk(metric, q, k)
        assert ix.stat("batches") == b0 + 1
        eng.set_option("no_batch", 1)
        try:
            loop = ix.scan_topk(metric, q, k)
        finally:
            eng.set_option("no_batch", 0)
        rowids = np.arange(1, n + 1, dtype=np.int64)
        for b in range(nq):
            d_b, d_l = res[b][1], loop[b][1]
            scale = np.maximum(np.abs(d_l), 1.0 if metric in (po.COS, po.DOT) else 1e-30)
            assert len(d_b) == len(d_l) and np.all(np.abs(d_b - d_l) <= 2e-5 * scale), (vtype, metric, b)
        for b in (0, nq - 1):   # and against the oracle (double / LASSQ accumulation in the reference)
            want_ids, want_d = oracle.scan_dense(metric, vtype, q[b], x, rowids, k)
            scale = np.maximum(np.abs(want_d), 1.0 if metric in (po.COS, po.DOT) else 1e-30)
            assert np.all(np.abs(res[b][1] - want_d) <= 2e-5 * scale)
            assert len(set(res[b][0].tolist()) ^ set(want_ids.tolist())) <= 2   # near ties may swap at the k-th place
        ix.close()
