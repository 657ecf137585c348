"""Property tests that pin the oracle to the UNMODIFIED reference build (oracle/_ref) on drawn inputs: distances for every
(type, metric) pair with special values mixed in, and the k-slot top-k (ties, start_max_index, k > n).  CPU only; skipped where
the reference tree was not available to build oracle/_ref (the committed golden fixtures cover those machines)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import pyoracle as po

TYPES = [po.F32, po.F16, po.BF16, po.U8, po.I8]
METRICS = [po.L2, po.L2SQ, po.COS, po.DOT, po.L1]
SPECIALS = [np.nan, np.inf, -np.inf, 0.0, -0.0, 65504.0, -65504.0, 1e-8, 3.0e38, -3.0e38, 5.96e-8]
FUZZ = settings(max_examples=400, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _same(a, b):
    a, b = np.float32(a), np.float32(b)
    return (np.isnan(a) and np.isnan(b)) or a.view(np.uint32) == b.view(np.uint32)


@FUZZ
@given(seed=st.integers(0, 2**32 - 1), dim=st.integers(1, 70), vtype=st.sampled_from(TYPES), metric=st.sampled_from(METRICS),
       nspecial=st.integers(0, 4), scale=st.sampled_from([1e-3, 1.0, 50.0, 3000.0]))
def test_distance_matches_reference(oracle, ref_cpu, seed, dim, vtype, metric, nspecial, scale):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = (rng.standard_normal(dim) * scale).astype(np.float32)
    b = (rng.standard_normal(dim) * scale).astype(np.float32)
    if vtype in (po.F32, po.F16, po.BF16):
        for _ in range(nspecial):
            (a if rng.integers(2) else b)[rng.integers(dim)] = rng.choice(SPECIALS)
    xa, xb = po.convert(a, vtype), po.convert(b, vtype)
    got, want = oracle.distance(metric, vtype, xa, xb), ref_cpu.distance(metric, vtype, xa, xb)
    assert _same(got, want), (vtype, metric, dim, a, b, got, want)


@FUZZ
@given(seed=st.integers(0, 2**32 - 1), n=st.integers(1, 400), dim=st.integers(1, 24), k=st.integers(1, 40), spread=st.integers(1, 6),
       qtype=st.sampled_from([po.Q_U8, po.Q_S8]), metric=st.sampled_from(METRICS), smi=st.integers(0, 39))
def test_quant_scan_topk_matches_reference(oracle, ref_cpu, seed, n, dim, k, spread, qtype, metric, smi):
    """vQuantRunMemory + vFullScanSortSlots on tie-heavy data, any start_max_index below k"""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi = (0, 2 * spread) if qtype == po.Q_U8 else (-spread, spread)
    vec = rng.integers(lo, hi + 1, (n, dim)).astype(np.uint8 if qtype == po.Q_U8 else np.int8)
    rowids = rng.permutation(n).astype(np.int64) * 7 - 3
    buf = np.zeros((n, 8 + dim), dtype=np.uint8)
    buf[:, :8] = rowids.view(np.uint8).reshape(n, 8)
    buf[:, 8:] = vec.view(np.uint8)
    q = rng.integers(lo, hi + 1, dim).astype(vec.dtype)
    smi = smi % k
    ids_o, d_o = oracle.scan_quant_buffer(metric, qtype, q, buf.reshape(-1), n, dim, k, start_max_index=smi)
    ids_r, d_r = ref_cpu.scan_quant_buffer(metric, qtype, q, buf.reshape(-1), n, dim, k, start_max_index=smi)
    assert np.array_equal(ids_o, ids_r) and np.array_equal(d_o, d_r), (n, dim, k, smi, qtype, metric)


@FUZZ
@given(seed=st.integers(0, 2**32 - 1), n=st.integers(1, 300), dim=st.integers(1, 20), k=st.integers(1, 30), vtype=st.sampled_from(TYPES),
       metric=st.sampled_from(METRICS))
def test_dense_scan_topk_matches_reference(oracle, ref_cpu, seed, n, dim, k, vtype, metric):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = po.convert(np.round(rng.standard_normal((n, dim)) * 2).astype(np.float32), vtype)       # coarse values: many equal distances
    rowids = np.arange(1, n + 1, dtype=np.int64)
    q = x[rng.integers(n)].copy()
    ids_o, d_o = oracle.scan_dense(metric, vtype, q, x, rowids, k)
    ids_r, d_r = ref_cpu.scan_dense(metric, vtype, q, x, rowids, k)
    assert np.array_equal(ids_o, ids_r) and np.array_equal(d_o, d_r), (vtype, metric, n, dim, k)
