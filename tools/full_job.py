"""The WHOLE BASELINE config on the GPUs of one box, through the C ABI's one-process multi-GPU entry (esfm_multi_*):
every pair of the triangle (configs[3]: 499,500 SURF pairs; configs[4]: 12,497,500 ORB pairs), bank replicated with one
ncclBroadcast, pairs dealt by work, every device's matches returned to this process chunk by chunk over its own PCIe link
(overlapped with the next chunk's sweep) -- "finishing the job" on the clock, not a slice of it and not an estimate.

  python tools/full_job.py surf --gpus 8                 # keeps all matches on the host (~16 GB)
  python tools/full_job.py orb  --gpus 8                 # keeps per-pair counts + 64-bit digests (the matches are ~2e10 records)
The printed JSON carries wall seconds including / excluding the bank broadcast, comparisons/s, per-device times and work
imbalance, and `output_sha1` = sha1(per-pair counts | per-pair match digests): equal at every --gpus iff the outputs are."""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

ap = argparse.ArgumentParser()
ap.add_argument("kind", choices=["surf", "orb"])
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--images", type=int, default=0)
ap.add_argument("--feat", type=int, default=0)
ap.add_argument("--keep", choices=["matches", "digests"], default=None)
ap.add_argument("--ratio", type=float, default=0.8)
args = ap.parse_args()
kind = args.kind
n_images = args.images or (1000 if kind == "surf" else 5000)
n_feat = args.feat or (8000 if kind == "surf" else 4000)
keep = args.keep or ("matches" if kind == "surf" else "digests")

multi = esfm.MultiContext(args.gpus)
mb = multi.bank(esfm.KIND_F32X64 if kind == "surf" else esfm.KIND_B256, n_images)
# the synthetic bank is produced ON device 0 (torch, seeded): declare the rows, let the library allocate, write into its buffer
for f in range(n_images):
    mb.primary.set_frame_rows(f, n_feat)
mb.primary.alloc_device()
ptr, nbytes = mb.primary.device_rows()
dev = torch.device("cuda:0")
raw = scheduler._wrap_device_bytes(ptr, nbytes, 0)
data = (synth.surf_like_torch if kind == "surf" else synth.orb_like_torch)(n_images, n_feat, 4 if kind == "surf" else 5, dev)
raw.copy_(data.reshape(-1).view(torch.uint8))
del data
torch.cuda.synchronize()

t0 = time.perf_counter()
mb.commit()                                   # device-0 commit + ncclBroadcast to the other devices
t_commit = time.perf_counter() - t0
tm_commit = multi.timing()
# warm-up on a corner of the triangle: derived layouts (built on a bank's first sweep), scratch and pinned-buffer growth
warm = mb.match_pairs(scheduler.all_pairs(min(n_images, 40)), args.ratio, True, keep=esfm.KEEP_DIGESTS)
warm.close()
st0 = multi.stats()
t1 = time.perf_counter()
res = mb.match_all_pairs(args.ratio, True, keep=esfm.KEEP_MATCHES if keep == "matches" else esfm.KEEP_DIGESTS)
t_match = time.perf_counter() - t1
tm = multi.timing()
st1 = multi.stats()
counts = res.pair_counts()
t2 = time.perf_counter()
digests = res.digests()
t_digest = time.perf_counter() - t2
h = hashlib.sha1()
h.update(counts.tobytes())
h.update(digests.tobytes())
n_pairs = res.n_pairs
comps = float(n_pairs) * n_feat * n_feat
per_dev = [{"pairs": int(b["pairs"] - a["pairs"]), "sweep_ms": b["sweep_ms_total"] - a["sweep_ms_total"],
            "sweep_launches": int(b["sweep_launches"] - a["sweep_launches"]), "d2h_bytes": int(b["d2h_bytes"] - a["d2h_bytes"])}
           for a, b in zip(st0, st1)]
sample = None
if keep == "matches":
    k = n_pairs // 2
    q, t, m = res.pair_at(k)
    sample = {"pair": [q, t], "matches": int(len(m)), "first": [int(m["queryIdx"][0]), int(m["trainIdx"][0]), float(m["distance"][0])] if len(m) else None}
print(json.dumps({
    "job": f"{kind} {n_images} x {n_feat} all pairs, ratio {args.ratio}, cross-check (BASELINE configs[{3 if kind == 'surf' else 4}])",
    "api": "esfm_multi_bank_commit + esfm_multi_match_all_pairs (one host process, one worker thread + stream per device)",
    "n_gpus": args.gpus, "pairs": int(n_pairs), "comparisons": comps, "keep": keep,
    "seconds_match_and_return": t_match, "seconds_bank_commit_and_broadcast": t_commit, "seconds_total": t_match + t_commit,
    "broadcast_ms": tm_commit["broadcast_ms"], "used_nccl": tm_commit["used_nccl"], "bank_bytes": int(nbytes),
    "comparisons_per_s": comps / t_match, "pairs_per_s": n_pairs / t_match,
    "comparisons_per_s_incl_broadcast": comps / (t_match + t_commit),
    "matches": int(res.n_matches), "match_bytes_returned_to_host": int(res.n_matches) * 16,
    "device_ms_max": tm["device_ms_max"], "device_ms_min": tm["device_ms_min"], "work_imbalance": tm["work_imbalance"],
    "per_device": per_dev, "digest_seconds_after_job": t_digest,
    "output_sha1": h.hexdigest(), "sample_pair": sample,
    "engines": {"l2": multi.contexts[0].l2_engine(), "hamming": multi.contexts[0].hamming_engine()},
}))
res.close()
mb.close()
multi.close()
