"""Scratch: throughput of fused group launches (one scan kernel for G independent queries) on one GPU."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vector_b200 as vs
from sqlite_vector_b200 import api
from tools.quick_bench import make_corpus

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_250_000); ap.add_argument("--dim", type=int, default=384); ap.add_argument("--queries", type=int, default=800)
a = ap.parse_args()
eng = vs.load_engine()
ix = vs.Index(api.I8, a.dim, a.n)
make_corpus(ix, api.I8, a.n, a.dim)
pitch = ix.query_pitch
qs = torch.randint(-60, 60, (64, pitch), dtype=torch.int8, device="cuda").view(torch.uint8)
qs[:, a.dim:] = 0
torch.cuda.synchronize()
st = torch.cuda.ExternalStream(ix.stream)
nslots = ix.stat("slots")
for G in (1, 2, 4, 8):
    def run(count):
        half = 0
        for g0 in range(0, count, G):
            ix.scan_submit_group(api.L2, qs[(g0 % 56)].data_ptr(), pitch, G, 20, True, half * 8, fetch=False)
            half ^= 1
    run(10 * G)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    run(a.queries)
    e1.record(st); e1.synchronize()
    us = e0.elapsed_time(e1) / a.queries * 1e3
    print(f"G={G}: {us:.1f} us/query -> {a.n*a.dim/us/1e3:.0f} GB/s, {1e6/us:.0f} qps")
