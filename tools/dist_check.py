"""torchrun --nproc-per-node N tools/dist_check.py [surf|orb]: the N-rank all-pairs result returned to rank 0 (work-balanced
deal, NCCL broadcast, chunked NCCL return) must be byte-identical to the single-GPU result (same library, same bank) -- also
with several chunk rounds per rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

kind = sys.argv[1] if len(sys.argv) > 1 else "orb"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ctx = esfm.Context(local)
rows = [900, 0, 1, 1300, 257, 1024, 640, 2, 777, 1500, 333, 1200]
frames = (synth.orb_like if kind == "orb" else synth.surf_like)(len(rows), rows, seed=31) if rank == 0 else None
shas = []
for chunk_env, reuse in ((None, False), ("6", True)):
    if chunk_env:
        os.environ["ESFM_CHUNK_PAIRS"] = chunk_env
    got = scheduler.match_all_pairs(frames, 0.8, True, ctx=ctx, block=4, reuse_staging=reuse)
    os.environ.pop("ESFM_CHUNK_PAIRS", None)
    if rank == 0:
        matches, off = got.all_matches()
        bank = ctx.bank_from_frames(frames)
        res = bank.match_all_pairs(0.8, True)
        ref, ref_off = res.all_matches()
        assert (res.pair_counts() == got.counts).all(), "per-pair counts differ"
        assert matches.tobytes() == ref.tobytes() and (off == ref_off).all(), "gathered matches differ from the single-GPU result"
        assert got.pair(9, 3).tobytes() == res.pair(9, 3).tobytes()
        shas.append(got.sha1())
        res.close(); bank.close()
if rank == 0:
    assert shas[0] == shas[1]
    print(f"dist_check[{kind}] world={world}: {got.n_pairs} pairs, {got.n_matches} matches, sha1={shas[0][:12]} identical to 1-GPU")
dist.barrier()
dist.destroy_process_group()
