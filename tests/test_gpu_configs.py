"""BASELINE.json config 5 in miniature (all 5 metrics x 5 dtypes, dims 128/384, n = 200k, batch 1; the stated size n = 1M x
dims {128,384,768,1536} runs in tests/test_gpu_at_size.py) plus the
recall of the int8-quantized scan against the exact f32 scan (the reference's own validation recipe,
QUANTIZATION.md:41-73).  -m gpu."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
TYPES = [po.F32, po.F16, po.BF16, po.U8, po.I8]
METRICS = [po.L2, po.L2SQ, po.COS, po.DOT, po.L1]


@pytest.mark.parametrize("dim", [128, 384])
@pytest.mark.parametrize("vtype", TYPES)
def test_config5_topk_all_metrics(oracle, vtype, dim):
    import sqlite_vector_b200 as vs
    n, k = 200_000, 20
    x = po.convert(po.gen_f32(n, dim, 1234), vtype)
    q = po.convert(po.gen_f32(1, dim, 4321), vtype)[0]
    rowids = np.arange(1, n + 1, dtype=np.int64)
    ix = vs.Index(vtype, dim, n)
    ix.append_dense(x)
    ix.finalize()
    for metric in METRICS:
        (res,) = ix.scan_topk(metric, q, k)
        want_ids, want_d = oracle.scan_dense(metric, vtype, q, x, rowids, k)
        if vtype in (po.U8, po.I8):
            assert np.array_equal(res[0], want_ids) and np.array_equal(res[1], want_d), (vtype, metric, dim)
        else:
            scale = np.maximum(np.abs(want_d), 1.0 if metric in (po.COS, po.DOT) else 1e-30)
            assert np.all(np.abs(res[1] - want_d) <= 1e-5 * scale), (vtype, metric, dim, np.abs(res[1] - want_d).max())
            recall = len(set(res[0].tolist()) & set(want_ids.tolist())) / k
            assert recall >= 0.95, (vtype, metric, dim, recall)      # only near-ties at the k-th place may differ
    ix.close()


def test_quantized_recall_vs_exact(oracle):
    """recall@20 of the int8 scan (reference quantization rule) against the exact f32 scan, both on the GPU"""
    import sqlite_vector_b200 as vs
    n, dim, k, nq = 100_000, 384, 20, 16
    xf = po.gen_f32(n, dim, 1234)
    qf = po.gen_f32(nq, dim, 4321)
    scale, offset, qt = oracle.quant_params(po.F32, xf)
    xq = np.stack([oracle.quantize(po.F32, xf[i], offset, scale, qt) for i in range(0, n)]) if n <= 1000 else None
    if xq is None:   # vectorised restatement of the S8 rule for speed (checked against the oracle on a slice)
        s = xf * np.float32(scale)
        xq = np.clip(np.trunc(np.where(s < 0, s - np.float32(0.5), s + np.float32(0.5))), -128, 127).astype(np.int8)
        assert np.array_equal(xq[:50], np.stack([oracle.quantize(po.F32, xf[i], offset, scale, qt) for i in range(50)]))
    ixf = vs.Index(po.F32, dim, n); ixf.append_dense(xf); ixf.finalize()
    ixq = vs.Index(po.I8, dim, n); ixq.append_dense(xq); ixq.finalize()
    hits = 0
    for b in range(nq):
        (ef,) = ixf.scan_topk(po.L2, qf[b], k)
        qq = oracle.quantize(po.F32, qf[b], offset, scale, qt)
        (eq,) = ixq.scan_topk(po.L2, qq, k)
        hits += len(set(ef[0].tolist()) & set(eq[0].tolist()))
    recall = hits / (nq * k)
    assert recall >= 0.90, recall
    ixf.close(); ixq.close()
