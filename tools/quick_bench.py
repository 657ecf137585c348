"""Scratch timing of the single-query scan on a synthetic resident corpus (device timing via CUDA events)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vector_b200 as vs  # noqa: E402
from sqlite_vector_b200 import api  # noqa: E402

TORCH_DT = {api.F32: torch.float32, api.F16: torch.float16, api.BF16: torch.bfloat16, api.U8: torch.uint8, api.I8: torch.int8}


def make_corpus(ix, vtype, n, dim, seed=1234, block=1 << 20):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for a in range(0, n, block):
        m = min(block, n - a)
        x = torch.randn((m, dim), generator=g, device="cuda", dtype=torch.float32)
        if vtype == api.I8:
            y = torch.clamp(torch.round(x * 24.0), -128, 127).to(torch.int8)
        elif vtype == api.U8:
            y = torch.clamp(torch.round(x.abs() * 48.0), 0, 255).to(torch.uint8)
        else:
            y = x.to(TORCH_DT[vtype])
        torch.cuda.synchronize()
        ix.append_device(y.data_ptr(), m)
    ix.finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--vtype", type=int, default=api.I8)
    ap.add_argument("--metric", type=int, default=api.L2)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--opts", type=str, default="")
    a = ap.parse_args()
    eng = vs.load_engine()
    print(eng.backend_name())
    ix = vs.Index(a.vtype, a.dim, a.n)
    t0 = time.time()
    make_corpus(ix, a.vtype, a.n, a.dim)
    print(f"corpus {a.n}x{a.dim} type {a.vtype} resident in {time.time() - t0:.1f}s")
    pitch = ix.query_pitch
    q = torch.zeros(pitch, dtype=torch.uint8, device="cuda")
    qsrc = torch.randn(a.dim, device="cuda")
    if a.vtype == api.I8:
        qv = torch.clamp(torch.round(qsrc * 24), -128, 127).to(torch.int8)
    elif a.vtype == api.U8:
        qv = torch.clamp(torch.round(qsrc.abs() * 48), 0, 255).to(torch.uint8)
    else:
        qv = qsrc.to(TORCH_DT[a.vtype])
    q[: qv.numel() * qv.element_size()] = qv.view(torch.uint8)
    torch.cuda.synchronize()
    st = torch.cuda.ExternalStream(ix.stream)
    bytes_per_query = a.n * a.dim * api.ELEM_SIZE[a.vtype]
    configs = [c for c in a.opts.split(";")] if a.opts else [""]
    for cfg in configs:
        for kv in filter(None, cfg.split(",")):
            name, val = kv.split("=")
            eng.set_option(name, int(val))
        for _ in range(3):
            ix.scan_device_query(a.metric, q.data_ptr(), a.k)
        ids, d = ix.collect_last(a.k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.iters):
            ix.scan_device_query(a.metric, q.data_ptr(), a.k)
        e1.record(st)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        t0 = time.perf_counter()
        qh = qv.view(torch.int16).cpu().numpy().view(np.uint16) if a.vtype in (api.F16, api.BF16) else qv.cpu().numpy()
        for _ in range(a.iters):
            ix.scan_topk(a.metric, qh, a.k)
        e2e = (time.perf_counter() - t0) / a.iters * 1e3
        print(f"[{cfg or 'default'}] device {ms:.3f} ms/query -> {bytes_per_query / ms / 1e6:.0f} GB/s ({1e3 / ms:.0f} qps); "
              f"host-to-host {e2e:.3f} ms ({1e3 / e2e:.0f} qps); top1 {ids[0]} {d[0]:.3f}")


if __name__ == "__main__":
    main()
