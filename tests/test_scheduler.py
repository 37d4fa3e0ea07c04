"""Host-side logic of the pair scheduler on CPU: pair order, block-cyclic sharding, and the variable-length gather
over torch.distributed with the gloo backend at world_size 2 (the N > 1 path of scheduler.gather_matches).
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


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from easysfm_b200 import scheduler as sch, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = synth.orb_like(7, [120, 0, 64, 200, 33, 150, 90], seed=3)   # every rank regenerates the same bank
    pairs = sch.all_pairs(len(frames))
    mine = sch.shard_pairs(len(pairs), rank, world, block=4)
    ms = [oracle.match(frames[i], frames[j], 0.8, True) for (i, j) in pairs[mine]]
    counts = np.array([len(m) for m in ms], np.int32)
    local = np.concatenate(ms) if ms else np.zeros(0, oracle.DMATCH_DTYPE)
    got = sch.gather_matches(mine, counts, local, len(pairs), None)
    if rank == 0:
        c, off, allm = got
        np.savez(os.path.join(tmp, "gathered.npz"), counts=c, offsets=off, matches=allm)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_matches_gloo_world2(tmp_path):
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
