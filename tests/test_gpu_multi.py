"""Several GPUs and the chunk pipeline.  Everything that does not need a second physical GPU runs on the driver's one-GPU box:
the chunk pipeline (several chunks per batch, pinned ring -> pageable segments, digests-only mode) and esfm_multi_* with the
SAME device listed twice (two contexts, two worker threads, deal + merge; replicas filled by a device copy instead of NCCL).
With >= 2 GPUs: esfm_multi_* over NCCL and the one-process-per-GPU scheduler (torchrun), both byte-identical to one GPU.
Reference loop served: cpp_code/test/sfm.cpp:140-161."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from easysfm_b200 import scheduler, synth
from util import assert_matches_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ROWS = [900, 0, 1, 1300, 257, 1024, 640, 2, 777, 1500, 333, 1200]


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("kind", ["orb", "surf"])
def test_chunk_pipeline_is_invisible(ctx, kind):
    """The same 66 pairs as one chunk and as 14 chunks of 5 (double-buffered arenas, match download overlapped with the next
    sweep, pinned ring -> pageable segments): byte-identical matches, counts, bulk copy and digests; digests-only batches
    keep the same digests without keeping matches."""
    import easysfm_b200 as esfm
    frames = (synth.orb_like if kind == "orb" else synth.surf_like)(len(ROWS), ROWS, seed=31)
    bank = ctx.bank_from_frames(frames)
    whole = bank.match_all_pairs(0.8, True)
    ref_all, ref_off = whole.all_matches()
    assert ref_all.tobytes() == b"".join(whole.pair_at(k)[2].tobytes() for k in range(whole.n_pairs))
    assert (np.diff(ref_off) == whole.pair_counts()).all()
    with _env(ESFM_CHUNK_PAIRS=5):
        assert bank.chunk_pairs() == 5
        parts = bank.match_all_pairs(0.8, True)
        dig = bank.match_pairs(scheduler.all_pairs(len(ROWS)), 0.8, True, keep=esfm.KEEP_DIGESTS)
    assert parts.n_pairs == whole.n_pairs == 66 and parts.n_matches == whole.n_matches == dig.n_matches
    got_all, got_off = parts.all_matches()
    assert got_all.tobytes() == ref_all.tobytes() and (got_off == ref_off).all()
    for k in (0, 7, 30, 65):
        assert parts.pair_at(k)[2].tobytes() == whole.pair_at(k)[2].tobytes()
    assert (parts.pair_counts() == whole.pair_counts()).all() and (dig.pair_counts() == whole.pair_counts()).all()
    assert (whole.digests() == parts.digests()).all() and (whole.digests() == dig.digests()).all()
    assert len(set(whole.digests().tolist())) > 30                       # (the 30 pairs with a 0/1/2-row frame share few digests)
    with pytest.raises(esfm.EsfmError):
        dig.pair_at(3)                                                   # matches were not kept
    i, j, m = whole.pair_at(40)
    assert_matches_equal(m, oracle.match(frames[i], frames[j], 0.8, True), exact_distance=(kind == "orb"))


@pytest.mark.parametrize("kind", ["orb", "surf"])
def test_multi_two_replicas_on_one_gpu(ctx, kind):
    """esfm_multi_* with device 0 listed twice: two contexts, two worker threads, the work-balanced deal and the merge --
    everything but NCCL -- must reproduce the single-context bytes; also with several chunks per replica and digests only."""
    import easysfm_b200 as esfm
    frames = (synth.orb_like if kind == "orb" else synth.surf_like)(len(ROWS), ROWS, seed=32)
    bank = ctx.bank_from_frames(frames)
    ref = bank.match_all_pairs(0.8, True)
    ref_all, ref_off = ref.all_matches()
    with esfm.MultiContext([0, 0]) as multi:
        multi.set_engines(l2=ctx.l2_engine(), hamming=ctx.hamming_engine())
        mb = multi.bank_from_frames(frames)
        t = multi.timing()
        assert t["used_nccl"] == 0 and t["commit_ms"] > 0
        for chunk in (0, 7):
            with _env(**({"ESFM_CHUNK_PAIRS": chunk} if chunk else {})):
                got = mb.match_all_pairs(0.8, True)
                dig = mb.match_all_pairs(0.8, True, keep=esfm.KEEP_DIGESTS)
            g_all, g_off = got.all_matches()
            assert g_all.tobytes() == ref_all.tobytes() and (g_off == ref_off).all()
            assert got.pair(9, 3).tobytes() == ref.pair(9, 3).tobytes()
            assert (dig.digests() == ref.digests()).all() and (dig.pair_counts() == ref.pair_counts()).all()
            got.close(); dig.close()
        sub = scheduler.all_pairs(len(ROWS))[5:50:3]
        a = mb.match_pairs(sub, 0.5, False)
        b = bank.match_pairs(sub, 0.5, False)
        assert a.all_matches()[0].tobytes() == b.all_matches()[0].tobytes()
        st = multi.stats()
        assert st[0]["pairs"] > 0 and st[1]["pairs"] > 0                 # both replicas worked
        assert multi.timing()["work_imbalance"] < 0.5
        a.close(); b.close(); mb.close()


def test_multi_over_nccl_two_gpus():
    """Two physical GPUs in one process: ncclCommInitAll + one ncclBroadcast of the bank; result bytes equal one GPU."""
    import torch
    import easysfm_b200 as esfm
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for kind in ("orb", "surf"):
        frames = (synth.orb_like if kind == "orb" else synth.surf_like)(len(ROWS), ROWS, seed=33)
        with esfm.Context(0) as c0:
            ref = c0.bank_from_frames(frames).match_all_pairs(0.8, True)
            ref_all, _ = ref.all_matches()
            ref_dig = ref.digests()
        with esfm.MultiContext(2) as multi:
            mb = multi.bank_from_frames(frames)
            assert multi.timing()["used_nccl"] == 1
            got = mb.match_all_pairs(0.8, True)
            assert got.all_matches()[0].tobytes() == ref_all.tobytes()
            assert (got.digests() == ref_dig).all()
            st = multi.stats()
            assert st[0]["pairs"] > 0 and st[1]["pairs"] > 0


@pytest.mark.parametrize("kind", ["orb", "surf"])
def test_two_ranks_match_single_gpu(kind):
    """One process per GPU (torchrun, NCCL broadcast + chunked NCCL return to rank 0) == single GPU, byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tools", "dist_check.py"), kind]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "identical to 1-GPU" in out.stdout


def test_per_pair_entry_reuses_its_bank(ctx):
    """esfm_match_descriptors (the unmodified caller's per-pair shape, sfm.cpp:153,156) keeps one two-frame bank per kind and
    refills it: growing, shrinking, empty and alternating kinds must all stay exact."""
    shapes = [(300, 500), (1500, 700), (64, 64), (2049, 3000), (0, 10), (10, 0), (700, 1500), (1, 2)]
    for n, (nq, nt) in enumerate(shapes):
        Qb, Tb = synth.orb_like(2, [nq, nt], seed=50 + n)
        assert_matches_equal(ctx.match_descriptors(Qb, Tb, 0.8, n % 2 == 0), oracle.match(Qb, Tb, 0.8, n % 2 == 0))
        Qs, Ts = synth.surf_like(2, [nq, nt], seed=70 + n)
        got, ref = ctx.match_descriptors(Qs, Ts, 0.8, True), oracle.match(Qs, Ts, 0.8, True)
        assert len(got) == len(ref) and (got["trainIdx"] == ref["trainIdx"]).all()
        # strided input (a cv::Mat ROI: step > row bytes)
        wide = np.zeros((nq, 48), np.uint8); wide[:, :32] = Qb
        assert_matches_equal(ctx.match_descriptors(wide[:, :32], Tb, 0.8, False), oracle.match(Qb, Tb, 0.8, False))


def test_bank_frames_set_out_of_order_and_twice(ctx):
    """Frames staged out of order, re-set, and mixed with page-locked sources still land where the bank wants them (the
    in-order fast path makes the upload mirror the bank itself; everything else is gathered on the device)."""
    import torch
    import easysfm_b200 as esfm
    frames = synth.orb_like(5, [300, 129, 512, 40, 700], seed=90)
    ref = ctx.bank_from_frames(frames).match_all_pairs(0.8, True).all_matches()[0].tobytes()
    b = ctx.bank(esfm.KIND_B256, 5)
    b.set_frame(3, frames[3])
    b.set_frame(0, frames[1])              # wrong data first ...
    b.set_frame(4, frames[4])
    pinned = torch.from_numpy(frames[2].copy()).pin_memory()
    b.set_frame_pinned(2, pinned.numpy())
    b.set_frame(1, frames[1])
    b.set_frame(0, frames[0])              # ... then set again
    b.commit()
    assert b.match_all_pairs(0.8, True).all_matches()[0].tobytes() == ref
    b2 = ctx.bank(esfm.KIND_B256, 2)
    b2.set_frame_rows(0, 10); b2.set_frame_rows(1, 10)
    b2.alloc_device()
    with pytest.raises(esfm.EsfmError):
        b2.commit()                        # rows declared without host data: esfm_bank_commit_device is the way


def test_hamming_frame_limit_is_a_whole_tile():
    """The XOR + POPC sweep keeps a frame's column minima in shared memory, sized by the TILE-PADDED row count: the bank must
    refuse exactly the frames the launch could not take (round-1 advisor: 49,793..49,904 rows were accepted, then failed)."""
    import easysfm_b200 as esfm
    from easysfm_b200 import synth
    with esfm.Context(0) as c:
        c.set_hamming_engine("popc")
        lim = None
        b = c.bank(esfm.KIND_B256, 1)
        for rows in range(49792, 49920, 16):
            try:
                b.set_frame_rows(0, rows)
                lim = rows
            except esfm.EsfmError:
                break
        assert lim is not None and lim % 128 == 0
        with pytest.raises(esfm.EsfmError):
            c.bank(esfm.KIND_B256, 1).set_frame(0, np.zeros((lim + 1, 32), np.uint8))
        big, small = synth.orb_like(2, [lim, 300], seed=77)
        got = c.match_descriptors(small, big, 0.8, True)            # train frame at the limit: must launch and be exact
        assert_matches_equal(got, oracle.match(small, big, 0.8, True))
