"""Scratch timing of the batched tensor-core path."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vector_b200 as vs
from sqlite_vector_b200 import api
from tools.quick_bench import make_corpus, TORCH_DT

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000); ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--vtype", type=int, default=api.I8); ap.add_argument("--metric", type=int, default=api.L2)
    ap.add_argument("--k", type=int, default=20); ap.add_argument("--nq", type=int, default=1024); ap.add_argument("--iters", type=int, default=3); ap.add_argument("--opts", type=str, default="")
    ap.add_argument("--sweep", type=str, default="", help="';'-separated option sets (name=value,...) timed one after the other on the same corpus")
    a = ap.parse_args()
    eng = vs.load_engine()
    for kv in filter(None, a.opts.split(",")):
        name, val = kv.split("="); eng.set_option(name, int(val))
    ix = vs.Index(a.vtype, a.dim, a.n)
    make_corpus(ix, a.vtype, a.n, a.dim)
    qs = torch.randn((a.nq, a.dim), device="cuda")
    if a.vtype == api.I8: q = torch.clamp(torch.round(qs * 24), -128, 127).to(torch.int8).cpu().numpy()
    elif a.vtype == api.U8: q = torch.clamp(torch.round(qs.abs() * 48), 0, 255).to(torch.uint8).cpu().numpy()
    elif a.vtype == api.BF16: q = qs.to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)
    else: q = qs.to(torch.float16).view(torch.int16).cpu().numpy().view(np.uint16)
    for cfg in (a.sweep.split(";") if a.sweep else [""]):
      for kv in filter(None, cfg.split(",")):
          name, val = kv.split("="); eng.set_option(name, int(val))
      if cfg: print(f"--- {cfg}")
      for it in range(a.iters):
        c0, k0 = ix.stat("batch_cands"), ix.stat("batch_kept")
        b0, u0 = ix.stat("batch_us"), ix.stat("tc_us")
        res = ix.scan_topk(a.metric, q, a.k)
        dt = (ix.stat("batch_us") - b0) * 1e-6
        tc = (ix.stat("tc_us") - u0) * 1e-3
        flops = 2.0 * a.nq * a.n * a.dim
        print(f"iter {it}: {dt*1e3:.2f} ms  -> {a.nq/dt:.0f} qps, {flops/dt/1e12:.1f} TOP/s; hits {ix.stat('batch_cands')-c0}, kept {ix.stat('batch_kept')-k0}, batches {ix.stat('batches')}, tc kernels {tc:.2f} ms")
    # spot check vs single path
    eng.set_option("no_batch", 1)
    one = ix.scan_topk(a.metric, q[:2], a.k)
    eng.set_option("no_batch", 0)
    print("match single-query path:", all(np.array_equal(res[b][0], one[b][0]) for b in range(2)))

if __name__ == "__main__":
    main()
