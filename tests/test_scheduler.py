"""Host-side logic of the pair scheduler on CPU: pair order, block-cyclic sharding, and the variable-length gather
over torch.distributed with the gloo backend at world_size 2 (the N > 1 path: scheduler.gather_round per chunk round).
The per-rank matcher here is the oracle -- this test is about the plumbing, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from easysfm_b200 import scheduler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_pairs_is_the_reference_loop_order():
    p = scheduler.all_pairs(5)
    ref = [(i, j) for i in range(5) for j in range(i)]     # cpp_code/test/sfm.cpp:140,143
    assert [tuple(x) for x in p] == ref
    assert all(scheduler.pair_index(i, j) == k for k, (i, j) in enumerate(ref))
    assert len(scheduler.all_pairs(1000)) == 499500 and len(scheduler.all_pairs(0)) == 0


@pytest.mark.parametrize("n,world,block", [(499500, 8, 64), (300, 2, 64), (55, 4, 8), (10, 8, 64), (0, 2, 4)])
def test_shards_partition_the_triangle(n, world, block):
    shards = [scheduler.shard_pairs(n, r, world, block) for r in range(world)]
    allidx = np.concatenate(shards) if shards else np.zeros(0)
    assert sorted(allidx.tolist()) == list(range(n))                    # disjoint cover
    if n >= world * block * 4:
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= block                          # balanced to one block


def test_pair_work_counts_each_pair_once():
    rows = [10, 0, 7]
    w = scheduler.pair_work(scheduler.all_pairs(3), rows)
    assert w.tolist() == [0, 70, 0]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_deal_pairs_balances_ragged_work():
    rows = [8000] * 3 + [100] * 40 + [5000] * 7
    pairs = scheduler.all_pairs(len(rows))
    work = scheduler.pair_work(pairs, rows)
    for world in (1, 2, 4, 8):
        owner = scheduler.deal_pairs(work, world, block=8)
        assert owner.min() >= 0 and owner.max() < world and len(owner) == len(pairs)
        loads = np.array([work[owner == r].sum() for r in range(world)], np.float64)
        assert loads.max() <= 1.15 * loads.mean() + 8 * 8000 * 8000          # within one block of the mean
        blocks = owner[: len(owner) // 8 * 8].reshape(-1, 8)
        assert (blocks == blocks[:, :1]).all()                                  # consecutive pairs stay together
    assert (scheduler.deal_pairs(work, 4, 8) == scheduler.deal_pairs(work, 4, 8)).all()


def test_sharded_results_table_and_gather():
    rng = np.random.default_rng(0)
    pairs = scheduler.all_pairs(9)
    per_pair = [np.zeros(int(rng.integers(0, 6)), scheduler.DMATCH_DTYPE) for _ in pairs]
    for k, m in enumerate(per_pair):
        m["queryIdx"] = np.arange(len(m)); m["trainIdx"] = k; m["distance"] = rng.random(len(m))
    res = scheduler.ShardedResults(pairs)
    order = rng.permutation(len(pairs))
    for ids in (np.sort(order[:10]), np.sort(order[10:11]), np.sort(order[11:])):      # three blobs, arbitrary pair subsets
        junk = np.zeros(3, scheduler.DMATCH_DTYPE)                                       # blobs need not be dense
        blob = np.concatenate([junk] + [per_pair[k] for k in ids])
        cnt = np.array([len(per_pair[k]) for k in ids], np.int32)
        off = 3 + np.concatenate([[0], np.cumsum(cnt)[:-1]])
        res.add(ids, cnt, off, blob)
    allm, off = res.all_matches()
    assert allm.tobytes() == np.concatenate(per_pair).tobytes()
    assert res.n_matches == sum(len(m) for m in per_pair)
    for k, (i, j) in enumerate(pairs):
        assert res.pair_at(k).tobytes() == per_pair[k].tobytes() == res.pair(i, j).tobytes()
        assert off[k + 1] - off[k] == len(per_pair[k])
    assert len(res.sha1()) == 40


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from easysfm_b200 import scheduler as sch, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = [120, 0, 64, 200, 33, 150, 90]
    frames = synth.orb_like(7, rows, seed=3)   # every rank regenerates the same bank
    pairs = sch.all_pairs(len(frames))
    owner = sch.deal_pairs(sch.pair_work(pairs, rows), world, block=4)
    mine = np.nonzero(owner == rank)[0]
    out = sch.ShardedResults(pairs) if rank == 0 else None
    chunk = 5                                   # several rounds; the last one is empty on one rank
    rounds = max((int((owner == r).sum()) + chunk - 1) // chunk for r in range(world))
    for c in range(rounds):
        ids = mine[c * chunk:(c + 1) * chunk]
        ms = [oracle.match(frames[i], frames[j], 0.8, True) for (i, j) in pairs[ids]]
        counts = np.array([len(m) for m in ms], np.int32)
        local = np.concatenate(ms) if ms else np.zeros(0, oracle.DMATCH_DTYPE)
        offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64) if len(ms) else np.zeros(0, np.int64)
        blob = torch.from_numpy(np.ascontiguousarray(local).view(np.int32).reshape(-1, 4).copy())
        got = sch.gather_round(ids, counts, offs, blob, None)
        if rank == 0:
            for g in got:
                out.add(*g)
        else:
            assert got is None
    if rank == 0:
        allm, off = out.all_matches()
        np.savez(os.path.join(tmp, "gathered.npz"), counts=out.counts, offsets=off[:-1], matches=allm)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rounds_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    import oracle
    from easysfm_b200 import synth
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    frames = synth.orb_like(7, [120, 0, 64, 200, 33, 150, 90], seed=3)
    pairs = scheduler.all_pairs(7)
    for k, (i, j) in enumerate(pairs):
        ref = oracle.match(frames[i], frames[j], 0.8, True)
        got = z["matches"][z["offsets"][k]: z["offsets"][k] + z["counts"][k]]
        assert len(got) == len(ref)
        assert (got["queryIdx"] == ref["queryIdx"]).all() and (got["trainIdx"] == ref["trainIdx"]).all()
        assert (got["distance"] == ref["distance"]).all()
