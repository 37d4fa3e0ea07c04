"""torchrun --nproc-per-node N tools/dist_check.py [surf|orb]: the N-rank all-pairs result gathered on rank 0 must be
byte-identical to the single-GPU result (same library, same bank)."""
import hashlib, os, sys
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
got = scheduler.match_all_pairs(frames, 0.8, True, ctx=ctx, block=4)
if rank == 0:
    pairs, counts, offsets, matches = got
    bank = ctx.bank_from_frames(frames)
    res = bank.match_all_pairs(0.8, True)
    ref = np.concatenate([res.pair_at(k)[2] for k in range(res.n_pairs)])
    assert (res.pair_counts() == counts).all(), "per-pair counts differ"
    assert matches.tobytes() == ref.tobytes(), "gathered matches differ from the single-GPU result"
    print(f"dist_check[{kind}] world={world}: {len(pairs)} pairs, {len(matches)} matches, sha1={hashlib.sha1(matches.tobytes()).hexdigest()[:12]} identical to 1-GPU")
dist.barrier()
dist.destroy_process_group()
