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
    """Fast path of the sharded scan: every rank launches its scan + filter on the engine stream, the fixed-size result
    blocks are all-gathered device-to-device (NCCL over NVLink) and one D2H copy hands them to vsb_merge_result_blocks.
    Requires implicit rowids (rowid = global row + 1), which is what the preloaded benchmark shards use."""

    def __init__(self, ix, engine, world: int, bounds, device):
        import torch
        self.torch, self.ix, self.eng, self.world = torch, ix, engine, world
        self.first_seq = np.asarray(bounds[:world], dtype=np.int64)
        self.views = []
        for slot in (0, 1):
            ptr, nbytes = None, None
            self.views.append(None)
        self.device = device
        self.estream = torch.cuda.ExternalStream(ix.stream, device=device)
        self.gathered = None
        self.host = None

    def _view(self, slot):
        if self.views[slot] is None:
            ptr, nbytes = self.ix.result_block(slot)
            self.views[slot] = self.torch.as_tensor(_DevView(ptr, nbytes), device=self.device)
            if self.gathered is None:
                self.gathered = self.torch.empty(self.world * nbytes, dtype=self.torch.uint8, device=self.device)
                self.host = self.torch.empty(self.world * nbytes, dtype=self.torch.uint8).pin_memory()
                self.stride = nbytes
        return self.views[slot]

    def submit(self, metric: int, d_query_ptr: int, k: int):
        """launch scan + filter of one query on the engine stream; returns a ticket for finish()"""
        slot = self.ix.scan_device_query(metric, d_query_ptr, k)
        return slot, self.estream.record_event(), k

    def finish(self, ticket):
        """all-gather the shards' result blocks of a submitted query and merge them (same result on every rank)"""
        import torch.distributed as dist
        torch = self.torch
        slot, ev, k = ticket
        blk = self._view(slot)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)                                 # only this query; a later submit may still be scanning
        dist.all_gather_into_tensor(self.gathered, blk)
        self.host.copy_(self.gathered, non_blocking=True)
        cur.synchronize()
        return self.eng.merge_result_blocks(self.host.numpy(), self.world, self.stride, self.first_seq, k)

    def query(self, metric: int, d_query_ptr: int, k: int):
        return self.finish(self.submit(metric, d_query_ptr, k))
