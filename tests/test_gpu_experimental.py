"""Parity of the EXPERIMENTAL engine options (compiled but off by default): only run when
VSB_TEST_EXPERIMENTAL=1, so that unvalidated variants can never turn the regular GPU suite red.  -m gpu."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VSB_TEST_EXPERIMENTAL") != "1", reason="set VSB_TEST_EXPERIMENTAL=1 to run")]


@pytest.mark.parametrize("option,value", [("epi_max", 1), ("tc_n", 128)])
@pytest.mark.parametrize("vtype", [po.I8, po.U8])
@pytest.mark.parametrize("metric", [po.L2, po.L2SQ, po.DOT, po.COS])
def test_batch_variant_bit_exact(oracle, option, value, vtype, metric):
    import sqlite_vector_b200 as vs
    eng = vs.load_engine()
    rng = np.random.Generator(np.random.PCG64(4000 + 10 * vtype + metric))
    n, dim, nq, k = 60000, 384, 300, 20
    x = po.convert(rng.standard_normal((n, dim), dtype=np.float32), vtype)
    q = po.convert(rng.standard_normal((nq, dim), dtype=np.float32), vtype)
    ix = vs.Index(vtype, dim, n)
    ix.append_dense(x)
    ix.finalize()
    old = eng.set_option(option, value)
    try:
        b0 = ix.stat("batches")
        res = ix.scan_topk(metric, q, k)
        assert ix.stat("batches") == b0 + 1, "the tensor-core batch path did not run"
    finally:
        eng.set_option(option, old)
    rowids = np.arange(1, n + 1, dtype=np.int64)
    for b in list(range(0, nq, 37)) + [nq - 1]:
        want_ids, want_d = oracle.scan_dense(metric, vtype, q[b], x, rowids, k)
        assert np.array_equal(res[b][0], want_ids) and np.array_equal(res[b][1], want_d), (option, vtype, metric, b)
    ix.close()
