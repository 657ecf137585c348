"""Oracle vs the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py).  Needs neither a GPU nor /root/reference."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TYPES = [po.F32, po.F16, po.BF16, po.U8, po.I8]
METRICS = [po.L2, po.L2SQ, po.COS, po.DOT, po.L1]


def _bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and np.array_equal(np.isnan(a), np.isnan(b))


@pytest.mark.parametrize("vt", TYPES)
def test_distances(oracle, vt):
    g = np.load(os.path.join(G, "distances.npz"))
    for dim in (5, 128, 384):
        x = g[f"x_{vt}_{dim}"]
        for m in METRICS:
            got = [oracle.distance(m, vt, x[0], x[i]) for i in range(25)]
            assert _bits_equal(got, g[f"d_{vt}_{dim}_{m}"]), (vt, dim, m)
            if vt in (po.U8, po.I8):
                got = [oracle.distance(m, vt, x[0], x[i], int_exact=True) for i in range(25)]
                assert _bits_equal(got, g[f"davx2_{vt}_{dim}_{m}"]), (vt, dim, m)


@pytest.mark.parametrize("vt", [po.F32, po.F16, po.BF16])
def test_special_values(oracle, vt):
    g = np.load(os.path.join(G, "distances.npz"))
    a, b = g[f"spa_{vt}"], g[f"spb_{vt}"]
    for m in METRICS:
        got = [oracle.distance(m, vt, a[i], b[i]) for i in range(40)]
        assert _bits_equal(got, g[f"spd_{vt}_{m}"]), (vt, m)


@pytest.mark.parametrize("vt", TYPES)
def test_quantize(oracle, vt):
    g = np.load(os.path.join(G, "quantize.npz"))
    x = g[f"x_{vt}"]
    for qt in (po.Q_U8, po.Q_S8):
        scale, offset = g[f"p_{vt}_{qt}"]
        got = np.stack([oracle.quantize(vt, x[r], offset, scale, qt).view(np.uint8) for r in range(8)])
        assert np.array_equal(got, g[f"q_{vt}_{qt}"])


def test_topk_quantized(oracle):
    g = np.load(os.path.join(G, "topk.npz"))
    for c in range(int(g["ncases"][0])):
        qt, n, dim, k = (int(v) for v in g[f"c{c}_meta"])
        vec, rowids, q = g[f"c{c}_vec"], g[f"c{c}_rowids"], g[f"c{c}_q"]
        buf = np.zeros((n, 8 + dim), dtype=np.uint8)
        buf[:, :8] = rowids.view(np.uint8).reshape(n, 8)
        buf[:, 8:] = vec.view(np.uint8)
        for m in METRICS:
            ids, d = oracle.scan_quant_buffer(m, qt, q, buf.reshape(-1), n, dim, k)
            assert np.array_equal(ids, g[f"c{c}_ids_{m}"]) and np.array_equal(d, g[f"c{c}_dist_{m}"]), (c, m)


@pytest.mark.parametrize("vt", [po.F32, po.F16, po.BF16])
def test_topk_fp(oracle, vt):
    g = np.load(os.path.join(G, "topk.npz"))
    x, q = g[f"fp{vt}_x"], g[f"fp{vt}_q"]
    rowids = np.arange(1, x.shape[0] + 1, dtype=np.int64)
    for m in METRICS:
        ids, d = oracle.scan_dense(m, vt, q, x, rowids, 20)
        assert np.array_equal(ids, g[f"fp{vt}_ids_{m}"]) and np.array_equal(d, g[f"fp{vt}_dist_{m}"]), (vt, m)
