#!/usr/bin/env python
"""bench.py — top-k queries/sec of the brute-force scan on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): vector_quantize_scan, int8 quantized from a synthetic N(0,1) f32
10M x 384 corpus with the reference's S8 rule, L2 (the default metric), k=20, batch=1.  A "step" is one query
scanned against the whole resident corpus (3.84 GB >> 126 MB L2, so every step streams from HBM).

Our arm prints `value` (queries already in HBM; launch -> candidates -> exact top-k on the host, per query) and
`e2e` (HOST query in, host top-k out through the C ABI: vsb_scan_submit/vsb_collect with two queries in flight, the
one-at-a-time vsb_scan_topk figure beside it; sharded runs: the grouped exchange fed with host queries).  `--impl reference` times the
reference's own CPU scan (oracle/_ref built from the unmodified sources with -mavx2 -mfma; falls back to the
oracle port) on all host cores.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_NAME = "top-k queries/sec, 10M x 384 f32 corpus (int8-quantized scan), k=20, batch=1"
SEED_CORPUS, SEED_QUERY = 1234, 4321
BLOCK = 1 << 19  # rows per generation block (block b is seeded with SEED_CORPUS + b, so shards are N-independent)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------ synthetic data (torch on the GPU: plumbing only)
def gen_block_f32(torch, b: int, rows: int, dim: int, device):
    g = torch.Generator(device=device).manual_seed(SEED_CORPUS + b)
    return torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)


def quantize_s8(torch, x, scale: float):
    """the reference's S8 rule: q = clamp(trunc(v*scale +- 0.5)) (src/sqlite-vector.c:626-656), offset 0"""
    s = x * scale
    r = torch.where(s < 0, s - 0.5, s + 0.5).trunc()
    return r.clamp_(-128, 127).to(torch.int8)


def corpus_blocks(n: int, lo: int, hi: int):
    """(block index, first row, rows) of the generation blocks overlapping rows [lo, hi)"""
    out = []
    for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        a0, a1 = b * BLOCK, min(n, (b + 1) * BLOCK)
        out.append((b, a0, a1 - a0))
    return out


def corpus_absmax(torch, n, dim, lo, hi, device):
    m = 0.0
    for b, a0, rows in corpus_blocks(n, lo, hi):
        x = gen_block_f32(torch, b, rows, dim, device)
        s, e = max(lo, a0) - a0, min(hi, a0 + rows) - a0
        m = max(m, float(x[s:e].abs().max()))
    return m


def make_queries(torch, count, dim, scale, device):
    g = torch.Generator(device=device).manual_seed(SEED_QUERY)
    qf = torch.randn((count, dim), generator=g, device=device, dtype=torch.float32)
    return quantize_s8(torch, qf, scale)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.first = 0

    def mark(self):
        """the timed legs start here; keep one earlier sample (taken under the warm-up load) in case the legs are short"""
        self.first = max(len(self.lines) - 1, 0)

    def wait_first(self, timeout_s: float):
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "nvidia-smi -lms 20 over the timed value leg and the timed e2e leg (plus the last sample before them)"}


# ------------------------------------------------------------------ CPU reference arm
def host_quant_buffer(x_i8: np.ndarray) -> np.ndarray:
    """the vector_quantize_preload layout n x [int64 LE rowid | dim bytes] (src/sqlite-vector.c:1295-1311)"""
    n, dim = x_i8.shape
    buf = np.empty((n, 8 + dim), dtype=np.uint8)
    buf[:, :8] = np.arange(1, n + 1, dtype=np.int64).view(np.uint8).reshape(n, 8)
    buf[:, 8:] = x_i8.view(np.uint8)
    return buf.reshape(-1)


def cpu_reference_run(buf, n, dim, k, queries_i8: np.ndarray, threads: int, reps: int):
    """times threads*reps queries of the reference's own vQuantRunMemory + vFullScanSortSlots; returns (seconds, kind)"""
    from oracle import pyoracle as po
    need = threads * reps
    q = np.ascontiguousarray(np.resize(queries_i8, (need, dim)))
    try:
        ref = po.RefHarness("avx2")
        assert ref.backend == "AVX2"
        sec = ref.time_queries(buf, n, dim, 8 + dim, 8, po.L2, po.I8, True, po.Q_S8, k, q, threads, reps)
        return sec, "reference"
    except (FileNotFoundError, OSError, AssertionError):
        orc = po.Oracle()

        def work(t):
            for r in range(reps):
                orc.scan_quant_buffer(po.L2, po.Q_S8, q[t * reps + r], buf, n, dim, k)
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        return time.perf_counter() - t0, "port"


# ------------------------------------------------------------------ main
_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line: route everything else (NCCL banners, library prints) to stderr"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU reference leg (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the batch=1024 tensor-core extras")
    ap.add_argument("--group", type=int, default=8, help="sharded runs: queries per exchange group (two groups in flight)")
    ap.add_argument("--engine-opt", action="append", default=[], help="name=value passed to vsb_set_option (experiments)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != a.gpus and world > 1:
        log(f"warning: WORLD_SIZE={world} but --gpus {a.gpus}; using WORLD_SIZE")
    n, dim, k, K, W = a.n, a.dim, a.k, a.steps, max(a.warmup, 0)
    workload = f"vector_quantize_scan int8 dim={dim} n={n} k={k} batch=1 L2"
    cores = a.cpu_threads or (os.cpu_count() or 1)

    import torch

    if a.impl == "reference":
        if rank != 0:
            return 0
        reference_arm(torch, a, n, dim, k, K, W, workload, cores)
        return 0

    import torch.distributed as dist

    import sqlite_vector_b200 as vs
    from sqlite_vector_b200 import api, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    eng = vs.load_engine()
    for kv in a.engine_opt:
        name, val = kv.split("=")
        eng.set_option(name, int(val))

    # ---- resident corpus shard (preload is outside the timed region, like the reference's preloaded buffer)
    bounds = shard.shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    t0 = time.perf_counter()
    amax = corpus_absmax(torch, n, dim, lo, hi, device)
    if world > 1:
        t = torch.tensor([amax], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        amax = float(t.item())
    scale = float(np.float32(127.0) / np.float32(amax))  # S8: 127/max(|min|,|max|), offset 0 (:1265-1268)
    ix = vs.Index(api.I8, dim, hi - lo, device=local_rank, first_seq=lo)
    host_parts = []
    want_host = (world == 1 and not a.no_cpu_baseline)
    for b, a0, rows in corpus_blocks(n, lo, hi):
        x = quantize_s8(torch, gen_block_f32(torch, b, rows, dim, device), scale)
        s, e = max(lo, a0) - a0, min(hi, a0 + rows) - a0
        xs = x[s:e].contiguous()
        torch.cuda.synchronize()
        ix.append_device(xs.data_ptr(), e - s)
        if want_host:
            host_parts.append(xs.cpu().numpy())
    ix.finalize()
    NQ = max(K + W + 8, 1024)
    q_all = make_queries(torch, NQ, dim, scale, device)                  # int8 [NQ, dim]
    pitch = ix.query_pitch
    q_dev = torch.zeros((NQ, pitch), dtype=torch.uint8, device=device)
    q_dev[:, :dim] = q_all.view(torch.uint8)
    q_host = q_all.cpu().numpy()
    torch.cuda.synchronize()
    log(f"[rank {rank}] shard rows [{lo},{hi}) resident in {time.perf_counter() - t0:.1f}s; scale={scale:.4f}")

    st = torch.cuda.ExternalStream(ix.stream, device=device)
    cap = 4096

    def step_device(i):
        ix.scan_device_query(api.L2, q_dev[i].data_ptr(), k)
        if world == 1:
            return ix.collect_last(k)
        # sharded: every rank scans its rows, candidates are all-gathered, the slot algorithm is replayed
        raise RuntimeError("unreachable")

    exch = shard.DeviceExchange(ix, eng, world, bounds, device, group=a.group) if world > 1 else None
    G = exch.group if exch else 1

    def step_sharded(i):
        # scan + filter on every rank, device-side all-gather of the result blocks, one D2H, C merge (slot replay)
        return exch.query(api.L2, q_dev[i].data_ptr(), k)

    def run_sharded(first, count, on_device):
        """`count` independent queries through the grouped exchange, two groups in flight: the all-gather + merge of
        group g overlaps the scans of group g+1.  on_device=False: host queries (pinned H2D inside the call)."""
        pending, last_res = None, None
        for g0 in range(0, count, G):
            i0, m = first + g0, min(G, count - g0)
            if on_device:
                t = exch.submit_strided(api.L2, q_dev[i0].data_ptr(), pitch, m, k, True)
            else:
                t = exch.submit_strided(api.L2, q_host[i0:i0 + m], q_host.strides[0], m, k, False)
            if pending is not None:
                last_res = exch.finish(pending)[-1]
            pending = t
        if pending is not None:
            last_res = exch.finish(pending)[-1]
        return last_res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: queries resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()         # nvidia-smi needs a moment before its first line: start it ahead of the warm-up
    run = step_device if world == 1 else step_sharded
    for i in range(W):
        run(i)
    if sampler:
        sampler.wait_first(2.0)
    eng.set_option("time_kernels", 4)      # every 4th query: event records between kernels cost host time and open small gaps
    ix.profile_read()
    barrier()
    if sampler:
        sampler.mark()          # samples from here on (value leg and e2e leg, both under the same load) are the ones reported
    l0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(st)
    last = None
    if world == 1:
        # independent single queries, two in flight: query i+1 is scanning while the host finishes query i
        pending = None
        for i in range(K):
            slot = ix.scan_device_query(api.L2, q_dev[W + i].data_ptr(), k)
            if pending is not None:
                last = ix.collect(pending, k)
            pending = slot
        last = ix.collect(pending, k)
    else:
        last = run_sharded(W, K, True)
    e1.record(st)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_dev = e0.elapsed_time(e1)
    launches = eng.kernel_launches() - l0
    prof = ix.profile_read()
    eng.set_option("time_kernels", 0)
    # the step time is the slower of the device-event span and the wall clock around the same region
    ms_total = max(ms_dev, t_wall * 1e3) if world > 1 else ms_dev
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # sequential single-query latency (launch -> complete exact top-k), for information
    lat_ms, lat_pct = None, None
    if world == 1:
        torch.cuda.synchronize()
        lats = []
        for i in range(min(K, 100)):
            t0l = time.perf_counter()
            step_device(W + i)
            lats.append((time.perf_counter() - t0l) * 1e3)
        lat_ms = float(np.mean(lats))
        lat_pct = {"p50": float(np.percentile(lats, 50)), "p95": float(np.percentile(lats, 95)), "n": len(lats)}

    # ---- e2e: host query in, host top-k out, through the public C-ABI call
    if world == 1:
        for i in range(W):
            ix.scan_topk(api.L2, q_host[i], k)
    else:
        run_sharded(0, W, False)
    barrier()
    sync_qps = None
    if world == 1:
        # (a) the synchronous call of the reference-facing plugin (what xFilter makes): one query at a time, each call
        #     = pinned H2D of the query + scan + filter + D2H of the candidate block + host slot replay
        t0s = time.perf_counter()
        for i in range(K):
            ix.scan_topk(api.L2, q_host[W + i], k)
        sync_qps = K / (time.perf_counter() - t0s)
    t0 = time.perf_counter()
    surv0, q0 = ix.stat("survivors"), ix.stat("queries")
    if world == 1:
        # (b) the asynchronous C-ABI pair vsb_scan_submit(host query) / vsb_collect with two queries in flight: the same
        #     copies per query, but query i+1 is staged and scanning while the host finishes query i
        pending = None
        for i in range(K):
            slot = ix.scan_submit(api.L2, q_host[W + i], k, on_device=False, fetch=True)
            if pending is not None:
                last_e2e = ix.collect(pending, k)
            pending = slot
        last_e2e = ix.collect(pending, k)
        assert np.array_equal(last_e2e[0], last[0]) and np.array_equal(last_e2e[1], last[1]), "host-query path != device-query path"
    else:
        last_e2e = run_sharded(W, K, False)
        assert np.array_equal(last_e2e[0], last[0]) and np.array_equal(last_e2e[1], last[1]), "host-query path != device-query path"
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    clocks = sampler.stop() if sampler else None
    avg_surv = (ix.stat("survivors") - surv0) / max(ix.stat("queries") - q0, 1) if world == 1 else None
    # header + block table + first 1024 candidate slots per shard, one cudaMemcpyAsync per query (group)
    d2h_bytes = int(ix.stat("fetch_bytes")) if world == 1 else int(exch.d2h_bytes_per_query)
    # ---- sharded runs: the batch=1024 half of the metric through the row-sharded tensor-core path (every rank takes part)
    batched_sharded = None
    if world > 1 and not a.no_batched:
        try:
            B = 1024
            shard.sharded_batch_topk(ix, api.L2, q_host[:B], k, bounds, device, as_arrays=True)   # warm-up: row norms, tensor maps
            barrier()
            reps, t0b = 3, time.perf_counter()
            for _ in range(reps):
                rb = shard.sharded_batch_topk(ix, api.L2, q_host[:B], k, bounds, device, as_arrays=True)
            barrier()
            dtb = (time.perf_counter() - t0b) / reps
            t = torch.tensor([dtb], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtb = float(t.item())
            if rb is None:
                batched_sharded = {"error": "the batch path refused on some shard"}
            else:
                batched_sharded = {"workload": f"vector_quantize_scan int8 dim={dim} n={n} k={k} batch={B} L2, {world} row shards",
                                   "queries_per_s": B / dtb, "ms_per_batch": dtb * 1e3, "batch": B,
                                   "end_to_end_tflops": 2.0 * dim * B * n / dtb / 1e12,
                                   "path": "per shard: tcgen05 scoring + exact refine + slot replay with entry logs; NCCL all-gather of the logs; GPU merge replay",
                                   "top1": [int(rb[0][0, 0]), float(rb[1][0, 0])]}
        except Exception as ex:  # never lose the headline line
            batched_sharded = {"error": str(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (scan_kernel): algorithmic bytes = rows * dim * 1 B per launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    scan_ms = prof["scan_ms"] / max(prof["scan_launches"], 1)
    shard_bytes = (hi - lo) * dim
    achieved = shard_bytes / (scan_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read + write of ONE scan launch from the newest committed `ncu --set full` capture of this workload
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*scan_kernel_traffic.json")))
    if tps and world == 1 and n == 10_000_000 and dim == 384:
        traffic = json.load(open(tps[-1])).get("dram_bytes_per_launch")

    out = {
        "metric": METRIC_NAME, "value": K / (ms_total * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8",
        "data": "synthetic N(0,1) f32 (seed 1234) quantized to int8 with the reference S8 rule; queries seed 4321",
        "config": {"workload": workload, "metric": "L2", "k": k, "batch": 1, "rows": n, "dim": dim, "shards": world,
                   "l2_flush": "none needed: each step streams the whole shard (%.2f GB) which exceeds the 126 MB L2" % (shard_bytes / 1e9),
                   "result_mode": "exact reference slot replay (bit-exact rowids/order/distances vs distance-cpu.c)",
                   "in_flight": 2 if world == 1 else 2 * G,
                   "exchange": None if world == 1 else f"groups of {G} independent queries: one NCCL all-gather of the shards' result blocks + one D2H per group, two groups in flight"},
        "e2e": {"value": K / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(pitch),
                "d2h_bytes_per_step": d2h_bytes, "avg_candidates_per_query": avg_surv,
                "synchronous_call_value": sync_qps,
                "note": ("vsb_scan_submit(host query)/vsb_collect, two queries in flight: per query a pinned H2D of the query, scan/filter kernels, one D2H copy of the candidate "
                         "block (header + table + 1024 slots), host slot replay; synchronous_call_value is the one-at-a-time vsb_scan_topk loop (the xFilter call)"
                         if world == 1 else
                         "vsb_scan_submit with HOST queries (pinned H2D per query) + scan/filter kernels on every rank + NCCL all-gather of the result blocks + D2H + host slot replay, same grouping as value")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "vsb::scan_kernel<int8,L2>", "avg_launch_ms": scan_ms, "launches_timed": prof["scan_launches"],
                     "filter_kernel_avg_ms": prof["filter_ms"] / max(prof["filter_launches"], 1), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(shard_bytes),
                     "note": "peak is the measured COPY bandwidth (reads + writes); this kernel only reads, so frac can exceed 1.0 "
                             "(ncu: dram__bytes_read = 1.0001 x algorithmic bytes per launch, profiles/r01e_ncu_summary.json); "
                             "avg_launch_ms is per query scan (a fused launch of G queries counts as G scans)"},
        "clocks": clocks,
        "wall_s_timed_region": t_wall,
        "single_query_latency_ms": lat_ms,
        "single_query_latency_percentiles_ms": lat_pct,
        "top1": {"rowid": int(last[0][0]), "distance": float(last[1][0])},
    }

    # ---- batched queries on the tensor cores (BASELINE metric "batch=1024"; configs[2]); informational extras
    if world == 1 and not a.no_batched:
        try:
            out["batched"] = batched_extras(torch, vs, api, ix, q_host, n, dim, k, device)
        except Exception as ex:  # never lose the headline line
            out["batched"] = {"error": str(ex)}

    if batched_sharded is not None:
        out["batched"] = {"int8_L2_dim%d_n%d_b1024_sharded" % (dim, n): batched_sharded}

    # ---- CPU baseline beside it (rank 0, N=1): the reference's AVX2 scan on a bounded sample
    if want_host:
        try:
            x_host = np.concatenate(host_parts)
            del host_parts
            buf = host_quant_buffer(x_host)
            qs = q_host[W:W + 8]
            sec1, kind = cpu_reference_run(buf, n, dim, k, qs, 1, 1)
            secN, kind = cpu_reference_run(buf, n, dim, k, qs, cores, 1)
            # parity spot-check of the timed GPU result against the reference on the same data
            from oracle import pyoracle as po
            chk = None
            try:
                ids_ref, d_ref = po.RefHarness("cpu").scan_quant_buffer(po.L2, po.Q_S8, q_host[W + K - 1], buf, n, dim, k)
                chk = bool(np.array_equal(ids_ref, last[0]) and np.array_equal(d_ref, last[1]))
            except (FileNotFoundError, OSError):
                ids_ref, d_ref = po.Oracle().scan_quant_buffer(po.L2, po.Q_S8, q_host[W + K - 1], buf, n, dim, k)
                chk = bool(np.array_equal(ids_ref, last[0]) and np.array_equal(d_ref, last[1]))
            out["cpu_baseline"] = {"value": cores / secN, "unit": "queries/s", "cores": cores, "kind": kind,
                                   "sample": f"{cores} independent queries (one per thread) over the same {n}x{dim} int8 preload buffer, "
                                             f"reference vQuantRunMemory+vFullScanSortSlots built -O3 -mavx2 -mfma; wall {secN:.2f}s",
                                   "single_thread_value": 1.0 / sec1, "single_thread_s_per_query": sec1,
                                   "gpu_result_matches_reference": chk}
        except MemoryError as ex:
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": cores, "kind": "reference", "sample": f"skipped: {ex}"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def batched_extras(torch, vs, api, ix, q_host, n, dim, k, device):
    """batch=1024 through vsb_scan_topk (host queries in, host top-k out): tcgen05 scoring + exact refinement"""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16_peak = float(json.load(open(peaks_path))["bf16_tflops"]) if os.path.exists(peaks_path) else 1590.0
    res = {}

    def run(name, index, queries, metric, flop_per_pair, peak, peak_note):
        B = queries.shape[0]
        index.scan_topk(metric, queries, k)                      # warm-up (row norms, tensor maps, workspaces)
        us0, rows0, bus0 = index.stat("tc_us"), index.stat("tc_rows"), index.stat("batch_us")
        reps = 10
        for _ in range(reps):
            r = index.scan_topk(metric, queries, k)
        dt = (index.stat("batch_us") - bus0) / reps * 1e-6       # wall time inside the C-ABI call (host queries in, host top-k out)
        tc_s = (index.stat("tc_us") - us0) / reps * 1e-6
        tc_rows = (index.stat("tc_rows") - rows0) / reps
        tc_tflops = flop_per_pair * B * tc_rows / tc_s / 1e12 if tc_s > 0 else None
        res[name] = {"queries_per_s": B / dt, "ms_per_batch": dt * 1e3, "batch": B,
                     "end_to_end_tflops": flop_per_pair * B * index.rows / dt / 1e12,
                     "roofline": {"bound": "tensor", "kernel": "vsb::tc_scan_kernel", "achieved": tc_tflops, "peak": peak, "unit": "TFLOP/s",
                                  "frac": (tc_tflops / peak) if tc_tflops else None, "peak_note": peak_note,
                                  "tc_kernel_ms_per_batch": tc_s * 1e3, "rows_scored_on_tensor_cores": tc_rows},
                     "top1": [int(r[0][0][0]), float(r[0][1][0])]}

    run("int8_L2_dim%d_n%d_b1024" % (dim, n), ix, q_host[:1024], api.L2, 2.0 * dim, 2 * bf16_peak,
        "int8 dense peak taken as 2x the measured bf16 cuBLAS peak (no measured int8 figure in MEASURED_PEAKS.json)")
    # BASELINE configs[2]: vector_full_scan dot bf16 dim=768 n=10M k=20 batch=1024
    n3, d3 = 10_000_000, 768
    ix3 = vs.Index(api.BF16, d3, n3, device=device.index or 0)
    for b, a0, rows in corpus_blocks(n3, 0, n3):
        x = gen_block_f32(torch, 7000 + b, rows, d3, device).to(torch.bfloat16)
        torch.cuda.synchronize()
        ix3.append_device(x.data_ptr(), rows)
    ix3.finalize()
    g = torch.Generator(device=device).manual_seed(SEED_QUERY + 1)
    q3 = torch.randn((1024, d3), generator=g, device=device).to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)
    run("bf16_DOT_dim768_n10M_b1024", ix3, q3, api.DOT, 2.0 * d3, bf16_peak, "measured cuBLAS bf16 burst peak (MEASURED_PEAKS.json bf16_tflops)")
    ix3.close()
    return res


def reference_arm(torch, a, n, dim, k, K, W, workload, cores):
    """the reference's own CPU scan on this box's host cores; each step = `cores` queries in parallel"""
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    amax = corpus_absmax(torch, n, dim, 0, n, device)
    scale = float(np.float32(127.0) / np.float32(amax))
    parts = []
    for b, a0, rows in corpus_blocks(n, 0, n):
        parts.append(quantize_s8(torch, gen_block_f32(torch, b, rows, dim, device), scale).cpu().numpy())
    buf = host_quant_buffer(np.concatenate(parts))
    del parts
    q = make_queries(torch, max(8, cores), dim, scale, device).cpu().numpy()
    # bounded sample: a step is `cores` independent queries (one per host thread) over the full corpus
    K = max(1, min(K, 3)); W = max(0, min(W, 1))
    kind = "reference"
    for _ in range(W):
        _, kind = cpu_reference_run(buf, n, dim, k, q, cores, 1)
    t = 0.0
    for _ in range(K):
        s, kind = cpu_reference_run(buf, n, dim, k, q, cores, 1)
        t += s
    val = K * cores / t
    out = {"impl": "reference", "metric": METRIC_NAME, "value": val, "unit": "queries/s", "n_gpus": a.gpus, "steps": K, "warmup": W,
           "ms_per_step": t / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8",
           "data": "synthetic N(0,1) f32 (seed 1234) quantized to int8 with the reference S8 rule; queries seed 4321",
           "config": {"workload": workload, "metric": "L2", "k": k, "batch": 1, "rows": n, "dim": dim,
                      "step": f"{cores} independent queries, one per host thread (bounded sample of the same workload)"},
           "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": kind,
                            "sample": f"{K} steps x {cores} queries over the {n}x{dim} preload buffer"},
           "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


if __name__ == "__main__":
    sys.exit(main())
