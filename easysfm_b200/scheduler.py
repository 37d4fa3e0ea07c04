"""The all-pairs loop of the reference as a sharded pair schedule, one process per GPU.

Reference: cpp_code/test/sfm.cpp:140-161 -- `for i in [0,N): for j in [0,i): match(frames[i], frames[j])`,
serial, one pair per call.  Here the N(N-1)/2 pairs are an explicit list that is
  * cut into blocks of consecutive pairs and dealt to the ranks of one box by WORK (rows_q * rows_t, so ragged frames
    do not unbalance the ranks) -- the same deterministic deal esfm_multi_match_pairs makes inside one process,
  * matched on each rank against a full replica of the descriptor bank (one NCCL broadcast from rank 0),
  * returned to rank 0 chunk by chunk: every rank's device-resident matches go to rank 0 with one NCCL send per chunk
    (counts and offsets first, then the 16-byte records), and rank 0 keeps them as per-rank blobs plus a per-pair
    (blob, offset, count) table -- no padded gather, no per-pair Python loop.
There is no inter-GPU traffic during matching: pairs are independent units.

Two ways to use several GPUs exist: this module (torchrun, one process per GPU: what bench.py --gpus N runs) and
esfm_multi_* / capi.MultiContext (ONE host process driving all GPUs, the shape of the reference's caller).

The torch.distributed plumbing is separated from the matching call so the host logic can be exercised on CPU with the
gloo backend (tests/test_scheduler.py, world_size 2).
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence

import numpy as np

from .capi import DMATCH_DTYPE, KIND_B256, KIND_F32X64

DEAL_BLOCK = 64


def all_pairs(n_frames: int) -> np.ndarray:
    """[(i, j)] with query = i > train = j, in the reference's loop order (sfm.cpp:140,143)."""
    i, j = np.tril_indices(n_frames, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)  # tril_indices is row-major: i ascending, j ascending within i


def pair_index(i: int, j: int) -> int:
    """Position of pair (i, j), j < i, in all_pairs order."""
    return i * (i - 1) // 2 + j


def shard_pairs(n_pairs: int, rank: int, world: int, block: int = DEAL_BLOCK) -> np.ndarray:
    """Indices (into the global pair list) owned by `rank`: blocks of `block` consecutive pairs dealt round-robin
    (equal-sized frames: the weak-scaling bench slices).  Consecutive pairs share the query frame, so a block keeps
    that frame hot in L2."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    idx = np.arange(n_pairs, dtype=np.int64)
    return idx[((idx // block) % world) == rank]


def pair_work(pairs: np.ndarray, rows: Sequence[int]) -> np.ndarray:
    """Comparisons per pair = rows_q * rows_t (counted once)."""
    r = np.asarray(rows, dtype=np.int64)
    return r[pairs[:, 0]] * r[pairs[:, 1]]


def deal_pairs(work: np.ndarray, world: int, block: Optional[int] = None) -> np.ndarray:
    """owner[pair]: blocks of `block` consecutive pairs, each to the rank with the least work so far (work + 1 per pair, as in
    csrc/multi.cu; default block: 64, smaller for small batches so that every rank still gets ~8 blocks).  Deterministic:
    every rank computes the same deal."""
    n = len(work)
    owner = np.zeros(n, np.int32)
    if world <= 1 or n == 0:
        return owner
    if block is None:
        block = max(1, min(DEAL_BLOCK, n // (world * 8)))
    nb = (n + block - 1) // block
    bw = np.add.reduceat(work.astype(np.float64) + 1.0, np.arange(0, n, block))
    load = np.zeros(world)
    bo = np.zeros(nb, np.int32)
    for b in range(nb):
        d = int(np.argmin(load))
        load[d] += bw[b]
        bo[b] = d
    return np.repeat(bo, block)[:n].astype(np.int32)


class ShardedResults:
    """Matches of a pair list held as blobs (one per rank and chunk) + a per-pair (blob, offset, count) table: what rank 0
    ends up with.  Nothing is copied pair by pair."""

    def __init__(self, pairs: np.ndarray):
        self.pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = len(self.pairs)
        self.counts = np.zeros(n, np.int32)
        self.offsets = np.zeros(n, np.int64)
        self.blob_id = np.full(n, -1, np.int32)
        self.blobs: List[np.ndarray] = []
        self._index = None

    def add(self, ids: np.ndarray, counts: np.ndarray, offsets: np.ndarray, blob: np.ndarray):
        self.blobs.append(blob)
        self.counts[ids] = counts
        self.offsets[ids] = offsets
        self.blob_id[ids] = len(self.blobs) - 1

    def from_results(self, res, ids: Optional[np.ndarray] = None):
        """Adopt a fetched capi.Results without copying: its host segments become blobs (and `res` stays open with self)."""
        segs, seg, off = res.segments()
        counts = res.pair_counts()
        ids = np.arange(res.n_pairs) if ids is None else np.asarray(ids)
        for s, blob in enumerate(segs):
            sel = seg == s
            if sel.any():
                self.add(ids[sel], counts[sel], off[sel], blob)
        self._keep = getattr(self, "_keep", []) + [res]

    @property
    def n_pairs(self) -> int:
        return len(self.pairs)

    @property
    def n_matches(self) -> int:
        return int(self.counts.sum(dtype=np.int64))

    def pair_at(self, k: int) -> np.ndarray:
        c = int(self.counts[k])
        if c == 0:
            return np.zeros(0, DMATCH_DTYPE)
        o = int(self.offsets[k])
        return self.blobs[int(self.blob_id[k])][o: o + c]

    def pair(self, query_frame: int, train_frame: int) -> np.ndarray:
        if self._index is None:
            self._index = {(int(q), int(t)): k for k, (q, t) in enumerate(self.pairs)}
        return self.pair_at(self._index[(int(query_frame), int(train_frame))])

    def all_matches(self):
        """(matches of every pair back to back in pair order, offsets[n_pairs + 1]) -- one vectorised gather per blob."""
        goff = np.zeros(self.n_pairs + 1, np.int64)
        np.cumsum(self.counts, dtype=np.int64, out=goff[1:])
        out = np.zeros(int(goff[-1]), DMATCH_DTYPE)
        for b, blob in enumerate(self.blobs):
            sel = np.nonzero((self.blob_id == b) & (self.counts > 0))[0]
            if len(sel) == 0:
                continue
            cnt = self.counts[sel].astype(np.int64)
            start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            within = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(start, cnt)
            out[np.repeat(goff[sel], cnt) + within] = blob[np.repeat(self.offsets[sel], cnt) + within]
        return out, goff

    def sha1(self) -> str:
        """Digest of (per-pair counts, all matches in pair order): equal across GPU counts iff the results are."""
        m, _ = self.all_matches()
        h = hashlib.sha1()
        h.update(self.counts.tobytes())
        h.update(m.tobytes())
        return h.hexdigest()


# ------------------------------------------------------------------------------------------------------
# torch.distributed plumbing (backend-agnostic: nccl on the GPU box, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------------
def gather_round(ids: np.ndarray, counts: np.ndarray, offsets: np.ndarray, blob, group=None, device=None, to_host=None):
    """One chunk round of the result return.  Every rank passes the global ids of the pairs it just matched, their match
    counts and offsets inside `blob` (a torch int32 tensor [n_matches, 4] on `device`: the 16-byte records).  Rank 0 gets
    [(ids, counts, offsets, blob ndarray)] for every rank with pairs in this round (its own included); other ranks get None.

    Wire protocol: one all_gather of (n_pairs, n_matches) per rank; then one send of the int64 table [n_pairs, 3] and one
    of the records per non-empty rank, received by rank 0 with batched irecvs."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    to_host = to_host or (lambda t: t.cpu().numpy())
    n_pairs, n_matches = len(ids), int(blob.shape[0]) if blob is not None else 0
    sizes = torch.tensor([n_pairs, n_matches], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [[int(x) for x in t.cpu().tolist()] for t in all_sizes]
    table = np.stack([np.asarray(ids, np.int64), np.asarray(counts, np.int64), np.asarray(offsets, np.int64)], axis=1) if n_pairs \
        else np.zeros((0, 3), np.int64)
    if rank != 0:
        ops = []
        if n_pairs:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(table)).to(dev), 0, group))
        if n_matches:
            ops.append(dist.P2POp(dist.isend, blob, 0, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return None
    tables, blobs, ops = {}, {}, []
    for r in range(1, world):
        k, m = all_sizes[r]
        if k:
            tables[r] = torch.empty((k, 3), dtype=torch.int64, device=dev)
            ops.append(dist.P2POp(dist.irecv, tables[r], r, group))
        if m:
            blobs[r] = torch.empty((m, 4), dtype=torch.int32, device=dev)
            ops.append(dist.P2POp(dist.irecv, blobs[r], r, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    out = []
    if n_pairs:
        own = to_host(blob).view(DMATCH_DTYPE).reshape(-1) if n_matches else np.zeros(0, DMATCH_DTYPE)
        out.append((table[:, 0], table[:, 1].astype(np.int32), table[:, 2], own))
    for r in range(1, world):
        if r in tables:
            t = tables[r].cpu().numpy()
            b = to_host(blobs[r]).view(DMATCH_DTYPE).reshape(-1) if r in blobs else np.zeros(0, DMATCH_DTYPE)
            out.append((t[:, 0], t[:, 1].astype(np.int32), t[:, 2], b))
    return out


def broadcast_bank(ctx, frames: Optional[Sequence[np.ndarray]], kind: Optional[int], group=None):
    """Replicate rank 0's descriptor bank on every rank's GPU: layout over the host channel, the descriptors with ONE NCCL
    broadcast straight into the library's device buffer (NVLink / NVSwitch)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    meta = [None]
    if rank == 0:
        meta[0] = (int(kind), [int(f.shape[0]) for f in frames])
    dist.broadcast_object_list(meta, src=0, group=group)
    kind, rows = meta[0]
    bank = ctx.bank(kind, len(rows))
    if rank == 0:
        for i, f in enumerate(frames):
            bank.set_frame(i, f)
        bank.commit()  # rank 0: host -> device upload (already in flight frame by frame since set_frame)
    else:
        for i, r in enumerate(rows):
            bank.set_frame_rows(i, r)
        bank.alloc_device()
    ptr, nbytes = bank.device_rows()
    if nbytes:
        t = _wrap_device_bytes(ptr, nbytes, ctx.device)
        dist.broadcast(t, src=0, group=group)
        torch.cuda.synchronize(ctx.device)
    if rank != 0:
        bank.commit_device()
    return bank, rows


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _wrap_device_bytes(ptr: int, nbytes: int, device: int):
    import torch
    with torch.cuda.device(device):
        return torch.as_tensor(_DevMem(ptr, nbytes), device=f"cuda:{device}")


class _PinnedStaging:
    """Grow-only pinned host buffers for the device->host leg on rank 0 (one per source rank), reused call after call."""

    def __init__(self):
        self.bufs = {}
        self.k = 0

    def begin(self):
        self.k = 0

    def __call__(self, t):
        import torch
        n = t.numel() * t.element_size()
        slot = self.k
        self.k += 1
        buf = self.bufs.get(slot)
        if buf is None or buf.numel() < n:
            buf = torch.empty((max(n + n // 4, 1 << 20),), dtype=torch.uint8, pin_memory=True)
            self.bufs[slot] = buf
        dst = buf[:n].view(t.dtype).reshape(t.shape)
        dst.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return dst.numpy()


_STAGING = _PinnedStaging()


def match_all_pairs(frames: Optional[Sequence[np.ndarray]], ratio: float, cross_check: bool, ctx=None, group=None,
                    block: Optional[int] = None, reuse_staging: bool = False, timing: Optional[dict] = None):
    """All pairs of `frames` (given on rank 0) across every rank of `group`.  Returns on rank 0 a ShardedResults in the
    reference's loop order; None on other ranks.  Without a process group: runs on ctx's GPU alone.
    reuse_staging=True keeps rank 0's copies of the other ranks' matches in pinned buffers that the NEXT call overwrites
    (bench loops); the default copies them into fresh memory."""
    import time

    import torch.distributed as dist

    from .capi import Context

    ctx = ctx if ctx is not None else Context(0)
    t0 = time.perf_counter()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        bank = ctx.bank_from_frames(frames)
        t1 = time.perf_counter()
        res = bank.match_all_pairs(ratio, cross_check)
        out = ShardedResults(all_pairs(len(frames)))
        out.from_results(res)      # zero-copy: the blobs alias the library's host segments, `out` keeps `res` alive
        bank.close()
        if timing is not None:
            timing.update(upload_broadcast_s=t1 - t0, match_gather_s=time.perf_counter() - t1)
        return out
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device(f"cuda:{ctx.device}")
    kind = None
    if rank == 0:
        kind = KIND_F32X64 if np.asarray(frames[0]).dtype == np.float32 else KIND_B256
    bank, rows = broadcast_bank(ctx, frames, kind, group)
    t1 = time.perf_counter()
    pairs = all_pairs(bank.n_frames)
    owner = deal_pairs(pair_work(pairs, rows), world, block)
    mine = np.nonzero(owner == rank)[0]
    chunk = bank.chunk_pairs() if len(pairs) else 1
    rounds = max((int((owner == r).sum()) + chunk - 1) // chunk for r in range(world)) if len(pairs) else 0
    out = ShardedResults(pairs) if rank == 0 else None
    if reuse_staging:
        _STAGING.begin()
    for c in range(rounds):
        ids = mine[c * chunk:(c + 1) * chunk]
        blob = torch.zeros((0, 4), dtype=torch.int32, device=dev)
        counts, offs = np.zeros(0, np.int32), np.zeros(0, np.int64)
        res = None
        if len(ids):
            res = bank.match_pairs(pairs[ids], ratio, cross_check, device_resident=True)
            counts = res.pair_counts()
            ptr, n, offs = res.device_matches()
            if n:
                blob = _wrap_device_bytes(ptr, n * DMATCH_DTYPE.itemsize, ctx.device).view(torch.int32).reshape(-1, 4)
        got = gather_round(ids, counts, offs, blob, group, device=dev, to_host=_STAGING if reuse_staging else None)
        if res is not None:
            res.close()
        if rank == 0:
            for (gi, gc, go, gb) in got:
                out.add(gi, gc, go, gb)
    bank.close()
    if timing is not None:
        timing.update(upload_broadcast_s=t1 - t0, match_gather_s=time.perf_counter() - t1)
    return out
