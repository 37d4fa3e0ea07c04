"""The whole BASELINE config on the GPUs of one box: every pair of the triangle, chunk by chunk, results fetched to the
host and folded into a checksum (so the device->host path is exercised) -- "finishing the job", not a slice of it.
usage: [torchrun ...] python tools/full_job.py [surf|orb] [n_images] [n_feat]"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

kind = sys.argv[1] if len(sys.argv) > 1 else "surf"
n_images = int(sys.argv[2]) if len(sys.argv) > 2 else (1000 if kind == "surf" else 5000)
n_feat = int(sys.argv[3]) if len(sys.argv) > 3 else (8000 if kind == "surf" else 4000)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = esfm.Context(local)
bank = ctx.bank(esfm.KIND_F32X64 if kind == "surf" else esfm.KIND_B256, n_images)
for f in range(n_images):
    bank.set_frame_rows(f, n_feat)
bank.alloc_device()
ptr, nbytes = bank.device_rows()
raw = scheduler._wrap_device_bytes(ptr, nbytes, local)
if rank == 0:
    data = (synth.surf_like_torch if kind == "surf" else synth.orb_like_torch)(n_images, n_feat, 4 if kind == "surf" else 5, dev)
    raw.copy_(data.reshape(-1).view(torch.uint8)); del data
torch.cuda.synchronize()
t0 = time.perf_counter()
if world > 1:
    dist.broadcast(raw, src=0); torch.cuda.synchronize()
t_bcast = time.perf_counter() - t0
bank.commit_device()
pairs = scheduler.all_pairs(n_images)
mine = pairs[scheduler.shard_pairs(len(pairs), rank, world, 64)]
CH = 16384
h = hashlib.sha1(); n_matches = 0
t1 = time.perf_counter()
for c0 in range(0, len(mine), CH):
    res = bank.match_pairs(mine[c0:c0 + CH], 0.8, True)
    n_matches += res.n_matches
    h.update(res.pair_counts().tobytes())
    res.close()
torch.cuda.synchronize()
t_match = time.perf_counter() - t1
tt = torch.tensor([t_match, float(n_matches)], dtype=torch.float64, device=dev)
if world > 1:
    tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    t_match = float(tmax[0]); n_matches = int(tt[1])
if rank == 0:
    comps = len(pairs) * n_feat * n_feat
    print(json.dumps({"job": f"{kind} {n_images} x {n_feat} all pairs, ratio 0.8, cross-check", "n_gpus": world, "pairs": int(len(pairs)),
                      "comparisons": comps, "seconds": t_match, "comparisons_per_s": comps / t_match, "pairs_per_s": len(pairs) / t_match,
                      "matches": n_matches, "bank_broadcast_s": t_bcast, "counts_sha1_rank0": h.hexdigest()[:16], "stats_rank0": ctx.stats()}))
if world > 1:
    dist.destroy_process_group()
