"""Multi-GPU plumbing: the flattened BFT is replicated on every GPU, the query batch is split into
contiguous equal ranges (sequences: on sequence boundaries) and only results travel — one gather to rank 0, or a sum
for the branching count. One process per GPU; `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU
tests) is the only collective layer. No exchange happens inside a lookup."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of items for `rank`; sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sequences(offs: np.ndarray, rank: int, world: int) -> Tuple[int, int]:
    """Split n sequences (offsets array of n+1 entries) so every rank gets about the same number of CHARACTERS,
    cutting only on sequence boundaries. Returns [begin, end) sequence indices."""
    n = len(offs) - 1
    if n <= 0:
        return 0, 0
    total = int(offs[-1]) - int(offs[0])
    lo = int(offs[0]) + (total * rank) // world
    hi = int(offs[0]) + (total * (rank + 1)) // world
    begin = int(np.searchsorted(offs[:-1], lo, side="left")) if rank else 0
    end = int(np.searchsorted(offs[:-1], hi, side="left")) if rank + 1 < world else n
    return begin, max(begin, end)


def gather_to_rank0(local, counts: List[int], group=None):
    """Gather per-rank result tensors (first dimension = this rank's item count) to rank 0, in rank order, i.e. in
    the original query order. Returns the concatenated tensor on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tail = tuple(local.shape[1:])
    m = max(counts) if counts else 0
    padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: counts[r]] for r in range(world)], dim=0)


def sum_to_all(value: int, device, group=None) -> int:
    """Sum of one integer over ranks (the -query_branching count)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


class ShardedQuery:
    """Runs one engine per rank on its shard and gathers results on rank 0. `engine` is a BFTEngine (GPU) or any
    object with the same query_* methods (the CPU tests pass a stub that wraps the oracle)."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _device(self):
        import torch
        d = getattr(self.engine, "device", None)
        return torch.device("cuda", d) if isinstance(d, int) and torch.cuda.is_available() else torch.device("cpu")

    def query_kmers(self, kmers: np.ndarray):
        """kmers: the FULL batch (same array on every rank). Returns (present, rows) for the full batch on rank 0."""
        import torch
        n = kmers.shape[0]
        b, e = shard_range(n, self.rank, self.world)
        present, rows, _ = self.engine.query_kmers(kmers[b:e])
        counts = [shard_range(n, r, self.world)[1] - shard_range(n, r, self.world)[0] for r in range(self.world)]
        dev = self._device()
        gp = gather_to_rank0(torch.from_numpy(present).to(dev), counts, self.group)
        gr = gather_to_rank0(torch.from_numpy(rows.view(np.int32)).to(dev), counts, self.group)
        if self.rank != 0:
            return None, None
        return gp.cpu().numpy(), gr.cpu().numpy().view(np.uint32)

    def query_sequences(self, chars: np.ndarray, offs: np.ndarray, threshold: float, canonical: bool):
        import torch
        offs = np.asarray(offs, dtype=np.uint64)
        ranges = [shard_sequences(offs, r, self.world) for r in range(self.world)]
        b, e = ranges[self.rank]
        sub_offs = offs[b:e + 1] - offs[b] if e > b else np.zeros(1, dtype=np.uint64)
        sub_chars = chars[int(offs[b]):int(offs[e])] if e > b else chars[:0]
        rows, _ = self.engine.query_sequences(sub_chars, sub_offs, threshold, canonical)
        counts = [r[1] - r[0] for r in ranges]
        g = gather_to_rank0(torch.from_numpy(rows.view(np.int32)).to(self._device()), counts, self.group)
        return None if self.rank != 0 else g.cpu().numpy().view(np.uint32)

    def query_branching_count(self, kmers: np.ndarray) -> int:
        n = kmers.shape[0]
        b, e = shard_range(n, self.rank, self.world)
        _, _, cnt = self.engine.query_branching(kmers[b:e])
        return sum_to_all(cnt, self._device(), self.group)


class PeerGather:
    """Gather fused into the query kernels: rank 0 owns the result arrays in its HBM, every other rank maps them
    through CUDA IPC (bft_b200_peer_import) and passes the slice for its shard as the kernel's OUTPUT pointer, so the
    result stores themselves cross NVLink/NVSwitch — no collective runs after the kernel. `torch.distributed` (any
    backend) is used only to ship the 64-byte handles and for the closing barrier."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _shared(self, nbytes: int):
        import torch.distributed as dist
        box = [None]
        base = None
        if self.rank == 0:
            base = self.engine.device_alloc(max(nbytes, 16))
            box[0] = self.engine.peer_export(base)
        dist.broadcast_object_list(box, src=0, group=self.group)
        ptr = base if self.rank == 0 else self.engine.peer_import(box[0])
        return base, ptr

    def _release(self, base, ptr):
        import torch.distributed as dist
        self.engine.sync()
        dist.barrier(group=self.group)          # every rank's stores have landed
        if self.rank != 0:
            self.engine.peer_close(ptr)

    def query_sequences(self, chars: np.ndarray, offs: np.ndarray, threshold: float, canonical: bool):
        """All ranks pass the FULL batch; returns the rows of every sequence on rank 0 (None elsewhere)."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        offs = np.asarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        base, ptr = self._shared(n * eng.RW * 4)
        b, e = shard_sequences(offs, self.rank, self.world)
        if e > b:
            dev = torch.device("cuda", eng.device)
            d_chars = torch.from_numpy(np.ascontiguousarray(chars[int(offs[b]):int(offs[e])])).to(dev)
            d_offs = torch.from_numpy((offs[b:e + 1] - offs[b]).astype(np.int64)).to(dev)
            eng.query_sequences_device(d_chars, d_offs, e - b, threshold, canonical, ptr + b * eng.RW * 4, None)
        self._release(base, ptr)
        out = None
        if self.rank == 0:
            out = eng.copy_from_device(base, np.empty((n, eng.RW), dtype=np.uint32))
            dist.barrier(group=self.group)
            eng.device_free(base)
        else:
            dist.barrier(group=self.group)
        return out

    def query_kmers(self, kmers: np.ndarray):
        """Presence bytes and colour rows of the full batch on rank 0, written there by every rank's kernel."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        n = kmers.shape[0]
        base_r, ptr_r = self._shared(n * eng.RW * 4)
        base_p, ptr_p = self._shared(n)
        b, e = shard_range(n, self.rank, self.world)
        if e > b:
            dev = torch.device("cuda", eng.device)
            d_k = torch.from_numpy(np.ascontiguousarray(kmers[b:e]).view(np.int64)).to(dev)
            eng.query_kmers_device(d_k, e - b, ptr_p + b, ptr_r + b * eng.RW * 4, None)
        self._release(base_r, ptr_r)
        if self.rank != 0:
            eng.peer_close(ptr_p)
        out = (None, None)
        if self.rank == 0:
            out = (eng.copy_from_device(base_p, np.empty(n, dtype=np.uint8)),
                   eng.copy_from_device(base_r, np.empty((n, eng.RW), dtype=np.uint32)))
            dist.barrier(group=self.group)
            eng.device_free(base_r)
            eng.device_free(base_p)
        else:
            dist.barrier(group=self.group)
        return out
