"""Scratch: per-CTA scan times and the effect of the adaptive row partition."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vector_b200 as vs
from sqlite_vector_b200 import api
from tools.quick_bench import make_corpus

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000); ap.add_argument("--dim", type=int, default=384); ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
eng = vs.load_engine()
ix = vs.Index(api.I8, a.dim, a.n)
make_corpus(ix, api.I8, a.n, a.dim)
pitch = ix.query_pitch
qs = torch.randint(-60, 60, (64, pitch), dtype=torch.int8, device="cuda").view(torch.uint8)
qs[:, a.dim:] = 0
torch.cuda.synchronize()
st = torch.cuda.ExternalStream(ix.stream)
nsm = torch.cuda.get_device_properties(0).multi_processor_count
for bal in (0, 1, 0, 1):
    eng.set_option("balance", bal)
    for i in range(30):                       # lets the partition settle
        ix.scan_device_query(api.L2, qs[i % 64].data_ptr(), 20)
    ix.collect_last(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(a.iters):
        ix.scan_device_query(api.L2, qs[i % 64].data_ptr(), 20)
    e1.record(st); e1.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    line = f"balance={bal}: {ms*1e3:.1f} us/query -> {a.n*a.dim/ms/1e6:.0f} GB/s"
    if bal or True:
        try:
            t = ix.debug_read("cta_time", np.uint32, nsm).astype(np.float64)
            b = ix.debug_read("bounds", np.int64, nsm + 1)
            sh = np.diff(b).astype(np.float64)
            line += f"; cta cycles min/mean/max {t.min():.0f}/{t.mean():.0f}/{t.max():.0f} (max/mean {t.max()/max(t.mean(),1):.3f}); shares min/max {sh.min():.0f}/{sh.max():.0f}"
        except Exception as ex:
            line += f"; ({ex})"
    print(line)
# is the imbalance systematic?  correlation of per-CTA times between consecutive queries with equal shares
eng.set_option("balance", 0)
ts = []
for i in range(6):
    ix.scan_device_query(api.L2, qs[i].data_ptr(), 20); ix.collect_last(20)
    try: ts.append(ix.debug_read("cta_time", np.uint32, nsm).astype(np.float64))
    except Exception: break
if len(ts) >= 2 and ts[0].size:
    c = np.corrcoef(np.stack(ts))
    print("per-CTA time correlation between queries (equal shares):", np.round(c[0, 1:], 3))
    order = np.argsort(ts[-1]); print("slowest CTAs:", order[-8:], "fastest:", order[:8])
