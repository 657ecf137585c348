"""f3 measurement: single-query scan of a STREAMED index (column in pinned host memory, two device windows): achieved
host->device GB/s and queries/s, next to the resident figure.  python tools/quick_stream.py [--n 10000000] [--window 1000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vector_b200 as vs  # noqa: E402
from sqlite_vector_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--window", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
rng = np.random.Generator(np.random.PCG64(1234))
ix = vs.Index(api.I8, a.dim, a.n, window_rows=a.window)
blk = 1 << 19
t0 = time.perf_counter()
for r0 in range(0, a.n, blk):
    m = min(blk, a.n - r0)
    ix.append_dense(rng.integers(-60, 61, (m, a.dim), dtype=np.int8))
ix.finalize()
fill_s = time.perf_counter() - t0
q = rng.integers(-60, 61, a.dim, dtype=np.int8)
ix.scan_topk(api.L2, q, 20)
b0, u0 = ix.stat("stream_bytes"), ix.stat("stream_us")
t0 = time.perf_counter()
for _ in range(a.iters):
    (res,) = ix.scan_topk(api.L2, q, 20)
dt = (time.perf_counter() - t0) / a.iters
gb = (ix.stat("stream_bytes") - b0) / a.iters / 1e9
print(json.dumps({"workload": f"streamed vector_quantize_scan int8 dim={a.dim} n={a.n} k=20, window {a.window} rows", "ms_per_query": dt * 1e3, "queries_per_s": 1 / dt,
                  "h2d_gb_per_query": gb, "h2d_gbs": gb / dt, "window_loop_gbs": (ix.stat("stream_bytes") - b0) / max(ix.stat("stream_us") - u0, 1) / 1e3,
                  "fill_s": fill_s, "top1": [int(res[0][0]), float(res[1][0])]}))
