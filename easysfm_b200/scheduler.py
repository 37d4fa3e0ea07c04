"""The all-pairs loop of the reference as a sharded pair schedule.

Reference: cpp_code/test/sfm.cpp:140-161 -- `for i in [0,N): for j in [0,i): match(frames[i], frames[j])`,
serial, one pair per call.  Here the N(N-1)/2 pairs are an explicit list that is
  * cut into blocks and dealt block-cyclically to the ranks of one box (one process per GPU),
  * matched on each rank against a full replica of the descriptor bank (NCCL broadcast from rank 0),
  * gathered back to rank 0 as compacted matches, in the reference's (i, j) order.
There is no inter-GPU traffic during matching: pairs are independent units.

The torch.distributed plumbing (broadcast / gather) is separated from the matching call so the host
logic can be exercised on CPU with the gloo backend (tests/test_scheduler.py, world_size 2).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .capi import DMATCH_DTYPE, KIND_B256, KIND_F32X64


def all_pairs(n_frames: int) -> np.ndarray:
    """[(i, j)] with query = i > train = j, in the reference's loop order (sfm.cpp:140,143)."""
    i, j = np.tril_indices(n_frames, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)  # tril_indices is row-major: i ascending, j ascending within i


def pair_index(i: int, j: int) -> int:
    """Position of pair (i, j), j < i, in all_pairs order."""
    return i * (i - 1) // 2 + j


def shard_pairs(n_pairs: int, rank: int, world: int, block: int = 64) -> np.ndarray:
    """Indices (into the global pair list) owned by `rank`: blocks of `block` consecutive pairs dealt round-robin.

    Consecutive pairs share the query frame, so a block keeps that frame hot in L2; round-robin dealing
    balances the triangle's work across ranks to within one block."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    idx = np.arange(n_pairs, dtype=np.int64)
    return idx[((idx // block) % world) == rank]


def pair_work(pairs: np.ndarray, rows: Sequence[int]) -> np.ndarray:
    """Comparisons per pair = rows_q * rows_t (counted once)."""
    r = np.asarray(rows, dtype=np.int64)
    return r[pairs[:, 0]] * r[pairs[:, 1]]


# ------------------------------------------------------------------------------------------------------
# torch.distributed plumbing (backend-agnostic: nccl on the GPU box, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------------
def gather_matches(local_pair_idx: np.ndarray, local_counts: np.ndarray, local_matches: np.ndarray, n_pairs: int,
                   group=None, device=None) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Variable-length gather of per-rank results to rank 0.

    local_pair_idx : int64 [k]   global indices of this rank's pairs (ascending)
    local_counts   : int32 [k]   matches per pair
    local_matches  : DMATCH_DTYPE [sum(counts)]  matches, pairs back to back in local_pair_idx order
    Returns on rank 0: (counts[n_pairs], offsets[n_pairs], matches) in global pair order; None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    # 1. everyone learns how many pairs / matches each rank holds
    sizes = torch.tensor([len(local_pair_idx), len(local_matches)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [t.cpu().numpy() for t in all_sizes]
    max_pairs = int(max(s[0] for s in all_sizes))
    max_matches = int(max(s[1] for s in all_sizes))
    # 2. padded gathers (pair ids + counts, then the 16-byte match records as int32 x 4)
    meta = torch.zeros((max(max_pairs, 1), 2), dtype=torch.int64, device=dev)
    if len(local_pair_idx):
        meta[: len(local_pair_idx), 0] = torch.from_numpy(np.ascontiguousarray(local_pair_idx, dtype=np.int64)).to(dev)
        meta[: len(local_pair_idx), 1] = torch.from_numpy(np.ascontiguousarray(local_counts, dtype=np.int64)).to(dev)
    rec = torch.zeros((max(max_matches, 1), 4), dtype=torch.int32, device=dev)
    if len(local_matches):
        rec[: len(local_matches)] = torch.from_numpy(
            np.ascontiguousarray(local_matches).view(np.int32).reshape(-1, 4)).to(dev)
    meta_list = [torch.zeros_like(meta) for _ in range(world)] if rank == 0 else None
    rec_list = [torch.zeros_like(rec) for _ in range(world)] if rank == 0 else None
    dist.gather(meta, meta_list, dst=0, group=group)
    dist.gather(rec, rec_list, dst=0, group=group)
    if rank != 0:
        return None
    counts = np.zeros(n_pairs, np.int32)
    src_rank = np.zeros(n_pairs, np.int32)
    src_off = np.zeros(n_pairs, np.int64)
    for r in range(world):
        k = int(all_sizes[r][0])
        m = meta_list[r][:k].cpu().numpy()
        ids, cnt = m[:, 0], m[:, 1]
        counts[ids] = cnt
        src_rank[ids] = r
        src_off[ids] = np.concatenate([[0], np.cumsum(cnt)[:-1]]) if k else np.zeros(0, np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)[:-1]]) if n_pairs else np.zeros(0, np.int64)
    out = np.zeros(int(counts.sum()), DMATCH_DTYPE)
    recs = [rec_list[r][: int(all_sizes[r][1])].cpu().numpy().view(DMATCH_DTYPE).reshape(-1) for r in range(world)]
    for p in range(n_pairs):
        c = counts[p]
        if c:
            out[offsets[p]: offsets[p] + c] = recs[src_rank[p]][src_off[p]: src_off[p] + c]
    return counts, offsets, out


def broadcast_bank(ctx, frames: Optional[Sequence[np.ndarray]], kind: Optional[int], group=None):
    """Replicate rank 0's descriptor bank on every rank's GPU: layout over the host channel, the 2 GB of
    descriptors with ONE NCCL broadcast straight into the library's device buffer (NVLink / NVSwitch)."""
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
        bank.commit()  # rank 0: host -> device upload + derived layouts
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
    return bank


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _wrap_device_bytes(ptr: int, nbytes: int, device: int):
    import torch
    with torch.cuda.device(device):
        return torch.as_tensor(_DevMem(ptr, nbytes), device=f"cuda:{device}")


def match_all_pairs(frames: Optional[Sequence[np.ndarray]], ratio: float, cross_check: bool, ctx=None, group=None,
                    block: int = 64):
    """All pairs of `frames` (given on rank 0) across every rank of `group`.  Returns on rank 0
    (pairs[n,2], counts[n], offsets[n], matches) in the reference's loop order; None on other ranks.
    Single-process use (no process group): runs on ctx's GPU alone."""
    import torch.distributed as dist

    from .capi import Context

    ctx = ctx if ctx is not None else Context(0)
    if not (dist.is_available() and dist.is_initialized()):
        bank = ctx.bank_from_frames(frames)
        res = bank.match_all_pairs(ratio, cross_check)
        pairs = all_pairs(len(frames))
        counts = res.pair_counts()
        ms = [res.pair_at(k)[2] for k in range(res.n_pairs)]
        offsets = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)[:-1]]) if len(counts) else np.zeros(0, np.int64)
        return pairs, counts, offsets, (np.concatenate(ms) if ms else np.zeros(0, DMATCH_DTYPE))
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    kind = None
    if rank == 0:
        kind = KIND_F32X64 if np.asarray(frames[0]).dtype == np.float32 else KIND_B256
    bank = broadcast_bank(ctx, frames, kind, group)
    pairs = all_pairs(bank.n_frames)
    mine = shard_pairs(len(pairs), rank, world, block)
    res = bank.match_pairs(pairs[mine], ratio, cross_check)
    counts = res.pair_counts()
    ms = [res.pair_at(k)[2] for k in range(res.n_pairs)]
    local = np.concatenate(ms) if ms else np.zeros(0, DMATCH_DTYPE)
    got = gather_matches(mine, counts, local, len(pairs), group, device=torch.device(f"cuda:{ctx.device}"))
    if got is None:
        return None
    return (pairs,) + got
