"""Row-sharded scan across ranks (one process per GPU, torch.distributed).

The corpus is split into contiguous row ranges in scan order (rank r owns rows
[n*r/W, n*(r+1)/W)), every rank scans its shard for the same query and emits its candidate list
(a small superset of the rows that can enter the reference's k slots, see include/vsb200.h), one
all-gather moves the fixed-size candidate blocks to every rank (NCCL over NVLink on GPUs, gloo in
the CPU tests) and the reference's slot algorithm is replayed over the concatenation in global scan
order (vsb_replay_topk) — which is exactly what the single-GPU path does with one shard.
"""
from __future__ import annotations

import numpy as np

from .api import CAND_DTYPE

REC = CAND_DTYPE.itemsize  # 24 bytes


def shard_bounds(n: int, world: int) -> list[int]:
    return [(n * r) // world for r in range(world + 1)]


def pack_candidates(cands: np.ndarray, cap: int) -> np.ndarray:
    """[int64 count][cap x 24-byte records] as one uint8 block of fixed size."""
    if cands.shape[0] > cap:
        raise ValueError(f"{cands.shape[0]} candidates exceed the gather capacity {cap}")
    blk = np.zeros(8 + cap * REC, dtype=np.uint8)
    blk[:8] = np.array([cands.shape[0]], dtype=np.int64).view(np.uint8)
    raw = np.ascontiguousarray(cands, dtype=CAND_DTYPE).view(np.uint8).reshape(-1)
    blk[8:8 + raw.size] = raw
    return blk


def unpack_candidates(blocks: np.ndarray, world: int, cap: int) -> np.ndarray:
    """inverse of pack_candidates for `world` concatenated blocks; result is in rank (= scan) order."""
    stride = 8 + cap * REC
    parts = []
    for r in range(world):
        blk = blocks[r * stride:(r + 1) * stride]
        cnt = int(blk[:8].view(np.int64)[0])
        parts.append(blk[8:8 + cnt * REC].view(CAND_DTYPE))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=CAND_DTYPE)


def allgather_candidates(cands: np.ndarray, cap: int, device=None) -> np.ndarray:
    """all-gather of the per-rank candidate blocks; returns the merged list (scan order) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    blk = torch.from_numpy(pack_candidates(cands, cap))
    if device is not None:
        blk = blk.to(device, non_blocking=True)
    out = torch.empty(world * blk.numel(), dtype=torch.uint8, device=blk.device)
    dist.all_gather_into_tensor(out, blk)
    return unpack_candidates(out.cpu().numpy(), world, cap)


def sharded_topk(engine, local_cands: np.ndarray, k: int, cap: int, device=None, max_index: int = 0):
    """merge step of one query: gather + replay.  Every rank returns the same (rowids, distances)."""
    merged = allgather_candidates(local_cands, cap, device)
    ids, d, mi = engine.replay_topk(merged, k, max_index)
    return ids, d, mi


class _DevView:
    """zero-copy torch view of engine-owned device memory (via __cuda_array_interface__)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class DeviceExchange:
    """Fast path of the sharded scan.  Every rank launches scan + filter of a GROUP of queries on the engine stream
    (one result slot each; the slots' result blocks are contiguous in device memory), the group's blocks are
    all-gathered device-to-device with ONE collective (NCCL over NVLink), one D2H copy brings them to pinned host
    memory and vsb_merge_result_blocks replays the reference's slot algorithm per query.  Groups are pipelined: while
    the blocks of group g are exchanged and merged, the GPU is already scanning group g+1.
    Requires implicit rowids (rowid = global row + 1), which is what the preloaded benchmark shards use."""

    def __init__(self, ix, engine, world: int, bounds, device, group: int = 4):
        import torch
        self.torch, self.ix, self.eng, self.world = torch, ix, engine, world
        self.first_seq = np.asarray(bounds[:world], dtype=np.int64)
        self.device = device
        self.estream = torch.cuda.ExternalStream(ix.stream, device=device)
        nslots = ix.stat("slots")
        self.group = max(1, min(group, nslots // 2))
        self.ngroups = nslots // self.group          # slot groups used round-robin; >= 2 so that two can be in flight
        base, nbytes = ix.result_block(0)
        self.stride = nbytes
        self.heads = torch.as_tensor(_DevView(base, nslots * nbytes), device=device)
        gbytes = world * self.group * nbytes
        self.gathered = [torch.empty(gbytes, dtype=torch.uint8, device=device) for _ in range(self.ngroups)]
        self.host = [torch.empty(gbytes, dtype=torch.uint8).pin_memory() for _ in range(self.ngroups)]
        self.next_group = 0

    def submit(self, metric: int, queries, k: int, on_device: bool = True):
        """launch scan + filter for up to `group` queries (a list of device pointers, or of host arrays when on_device is
        False); returns a ticket for finish()"""
        nq = len(queries)
        if not 0 < nq <= self.group:
            raise ValueError(f"a group holds 1..{self.group} queries")
        gi = self.next_group
        self.next_group = (gi + 1) % self.ngroups
        for j, q in enumerate(queries):
            self.ix.scan_submit(metric, q, k, on_device=on_device, fetch=False, slot=gi * self.group + j)
        return gi, nq, self.estream.record_event(), k

    def submit_strided(self, metric: int, queries, stride: int, nq: int, k: int, on_device: bool = True):
        """the same with ONE engine call: query j at `queries` + j * stride bytes (device pointer, or a contiguous host array)"""
        if not 0 < nq <= self.group:
            raise ValueError(f"a group holds 1..{self.group} queries")
        gi = self.next_group
        self.next_group = (gi + 1) % self.ngroups
        self.ix.scan_submit_group(metric, queries, stride, nq, k, on_device, gi * self.group)
        return gi, nq, self.estream.record_event(), k

    def finish(self, ticket):
        """all-gather the shards' result blocks of a submitted group and merge them; returns one (rowids, distances)
        pair per query, the same on every rank"""
        import torch.distributed as dist
        torch = self.torch
        gi, nq, ev, k = ticket
        lo = gi * self.group * self.stride
        blk = self.heads[lo:lo + nq * self.stride]
        out = self.gathered[gi][:self.world * nq * self.stride]
        host = self.host[gi][:self.world * nq * self.stride]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)                                 # only this group; a later group may still be scanning
        dist.all_gather_into_tensor(out, blk)              # rank r's blocks land at r * nq * stride
        host.copy_(out, non_blocking=True)
        cur.synchronize()
        return self.eng.merge_result_groups(host.numpy(), self.world, nq * self.stride, self.stride, nq, self.first_seq, k)

    def query(self, metric: int, query, k: int, on_device: bool = True):
        return self.finish(self.submit(metric, [query], k, on_device))[0]

    @property
    def d2h_bytes_per_query(self) -> int:
        return self.world * self.stride


class PeerExchange:
    """The sharded single-query path of round 2: no collective library on the data path.  Setup all-gathers one 64-byte
    cudaIpc handle per rank (torch.distributed object gather: plumbing); afterwards every filter kernel stores its result
    head straight into every peer's gather buffer over NVLink and the engine waits / copies / merges (vsb_exchange_*).
    Same interface as DeviceExchange (submit_strided / finish); requires implicit rowids (rowid = global row + 1)."""

    def __init__(self, ix, engine, world: int, rank: int, bounds, group: int = 8):
        import torch.distributed as dist
        self.ix, self.eng, self.world, self.rank = ix, engine, world, rank
        self.first_seq = np.asarray(bounds[:world], dtype=np.int64)
        nslots = ix.stat("slots")
        self.group = max(1, min(group, 8, nslots // 4))
        self.ngroups = nslots // self.group          # slot groups used round-robin; at most ngroups // 2 may be in flight
        self.max_in_flight = self.ngroups // 2
        mine = ix.exchange_export(world, rank)
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        ix.exchange_attach(b"".join(handles))
        dist.barrier()                               # every rank has mapped every buffer before the first push
        self.next_group = 0
        self.stride = ix.stat("fetch_bytes")

    def submit_strided(self, metric: int, queries, stride: int, nq: int, k: int, on_device: bool = True):
        if not 0 < nq <= self.group:
            raise ValueError(f"a group holds 1..{self.group} queries")
        gi = self.next_group
        self.next_group = (gi + 1) % self.ngroups
        self.ix.exchange_submit(metric, queries, stride, nq, k, on_device, gi * self.group)
        return gi, nq, k

    def finish(self, ticket):
        gi, nq, k = ticket
        return self.ix.exchange_collect(gi * self.group, nq, self.first_seq, k)

    def query(self, metric: int, query, k: int, on_device: bool = True):
        q = query if on_device else np.ascontiguousarray(query).reshape(1, -1)
        return self.finish(self.submit_strided(metric, q, 0, 1, k, on_device))[0]

    @property
    def d2h_bytes_per_query(self) -> int:
        return self.world * self.stride

    # ---- batched queries over the same exchange (implicit rowids: rowid = global row + 1)
    def batch_submit(self, metric: int, queries: np.ndarray, k: int):
        """tensor-core levels on this shard + push of the entry logs to every peer + device-side wait + GPU merge, all enqueued;
        returns a ticket for batch_finish (two batches may be in flight).  Raises VsbError(rc = ERANGE) when the batch path does
        not apply: use the per-query exchange then."""
        q2 = np.ascontiguousarray(queries).reshape(-1, queries.shape[-1])
        return self.ix.exchange_batch_submit(metric, q2, k, self.first_seq), q2.shape[0], k

    def batch_finish(self, ticket, as_arrays: bool = False):
        t, nq, k = ticket
        seq, d, counts = self.ix.exchange_batch_collect(t, nq, k)
        ids = seq + 1
        if as_arrays:
            return ids, d, counts
        return [(ids[b, :counts[b]].copy(), d[b, :counts[b]].copy()) for b in range(nq)]


def sharded_batch_topk(ix, metric: int, queries: np.ndarray, k: int, bounds, device, implicit_rowids: bool = True, as_arrays: bool = False):
    """Batched queries over a row-sharded column (BASELINE config 4).  Every rank runs the tensor-core batch path over
    its shard (vsb_batch_shard_scan), the shards' entry-log blocks are all-gathered device-to-device (NCCL over NVLink)
    and every rank replays them in shard order on its GPU (vsb_batch_merge) — bit for bit the result of one scan over the
    whole column.  Returns one (rowids, distances) pair per query — or, with as_arrays, (rowids[nq, k], distances[nq, k],
    counts[nq]) without the per-query Python objects — or None on every rank when the batch path does not apply to some
    shard (the caller then uses the per-query exchange)."""
    import torch
    import torch.distributed as dist

    from .api import ERANGE, VsbError

    world = dist.get_world_size()
    q2 = np.ascontiguousarray(queries).reshape(-1, queries.shape[-1])
    nq = q2.shape[0]
    ptr = nbytes = 0
    try:
        ptr, nbytes = ix.batch_shard_scan(metric, q2, k)
        bad = 0
    except VsbError as ex:
        if ex.rc != ERANGE:      # only "the batch path does not apply / a capacity was exceeded" is a reason to fall back
            raise
        bad = 1
    # a shard that cannot take the batch path must not leave the others waiting in the collective
    flag = torch.tensor([bad], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        return None
    blk = torch.as_tensor(_DevView(ptr, nbytes), device=device)
    gathered = torch.empty(world * nbytes, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, blk)
    torch.cuda.current_stream(device).synchronize()
    try:
        seq, d, counts = ix.batch_merge(gathered.data_ptr(), world, nbytes, np.asarray(bounds[:world], dtype=np.int64), nq, k)
    except VsbError as ex:
        if ex.rc != ERANGE:
            raise
        return None          # an entry log overflowed: the same verdict on every rank (same gathered data)
    if implicit_rowids:
        ids = seq + 1
    else:
        t = torch.from_numpy(ix.lookup_rowids(seq)).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ids = t.cpu().numpy()
    if as_arrays:
        return ids, d, counts
    return [(ids[b, :counts[b]].copy(), d[b, :counts[b]].copy()) for b in range(nq)]


def query_split(nq: int, world: int) -> list[int]:
    """contiguous query ranges of a query-sharded batch: rank r answers queries [b[r], b[r+1])"""
    return [(nq * r) // world for r in range(world + 1)]


def query_sharded_batch_topk(ix, metric: int, queries: np.ndarray, k: int, device=None, gather: bool = True):
    """Batched queries over REPLICAS of a column that fits one GPU: every rank holds the whole column (`ix` = the full index on
    this rank's GPU), the batch is split by query (query_split), every rank runs the tensor-core batch path on its slice
    (vsb_scan_topk) — no exchange on the data path, throughput scales with the number of GPUs.  This is the layout to use for
    batch throughput whenever n * dim * element size fits the HBM of one GPU (10M x 384 int8 = 3.84 GB; config 4 = 77 GB);
    row shards (PeerExchange / sharded_batch_topk) are for single-query latency and for columns larger than one GPU.

    Returns (rowids[nq, k], distances[nq, k], counts[nq]) for ALL queries on every rank when `gather` (one all-gather of the
    results: k * 12 + 4 bytes per query), else for this rank's slice only together with its (lo, hi)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    q2 = np.ascontiguousarray(queries).reshape(-1, queries.shape[-1])
    nq = q2.shape[0]
    b = query_split(nq, world)
    lo, hi = b[rank], b[rank + 1]
    if hi > lo:
        ids, dd, cnt = ix.scan_topk(metric, q2[lo:hi], k, as_arrays=True)
    else:
        ids, dd, cnt = np.zeros((0, k), dtype=np.int64), np.zeros((0, k), dtype=np.float64), np.zeros(0, dtype=np.int32)
    if not gather:
        return (ids, dd, cnt), (lo, hi)
    per = max(b[r + 1] - b[r] for r in range(world))
    blk = np.zeros((per, 2 * k + 1), dtype=np.int64)            # [rowids | distances as bit patterns | count], padded to the longest slice
    blk[:hi - lo, :k] = ids
    blk[:hi - lo, k:2 * k] = dd.view(np.int64)
    blk[:hi - lo, 2 * k] = cnt
    t = torch.from_numpy(blk.reshape(-1))
    if device is not None:
        t = t.to(device, non_blocking=True)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    o = out.cpu().numpy().reshape(world, per, 2 * k + 1)
    all_ids = np.concatenate([o[r, :b[r + 1] - b[r], :k] for r in range(world)])
    all_d = np.concatenate([o[r, :b[r + 1] - b[r], k:2 * k] for r in range(world)]).view(np.float64)
    all_c = np.concatenate([o[r, :b[r + 1] - b[r], 2 * k] for r in range(world)]).astype(np.int32)
    return all_ids, all_d, all_c
