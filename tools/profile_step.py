"""One device-resident match step per descriptor kind, sized for profiling under ncu (short kernels).
usage: python tools/profile_step.py [surf|orb] [n_frames] [n_feat] [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
from easysfm_b200 import synth, scheduler

kind = sys.argv[1] if len(sys.argv) > 1 else "surf"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 25
n_feat = int(sys.argv[3]) if len(sys.argv) > 3 else (8000 if kind == "surf" else 4000)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
ctx = esfm.Context(0)
data = (synth.surf_like_torch if kind == "surf" else synth.orb_like_torch)(n_frames, n_feat, 4, dev)
bank = ctx.bank(esfm.KIND_F32X64 if kind == "surf" else esfm.KIND_B256, n_frames)
for f in range(n_frames):
    bank.set_frame_rows(f, n_feat)
bank.alloc_device()
ptr, nbytes = bank.device_rows()
scheduler._wrap_device_bytes(ptr, nbytes, 0).copy_(data.reshape(-1).view(torch.uint8))
torch.cuda.synchronize()
bank.commit_device()
pairs = scheduler.all_pairs(n_frames)
for r in range(reps):
    s0 = ctx.stats()
    res = bank.match_pairs(pairs, 0.8, True, device_resident=True)
    s1 = ctx.stats()
    comps = s1["comparisons"] - s0["comparisons"]
    ms = s1["last_sweep_ms"]
    ops = 128 if kind == "surf" else 8
    peak = 148 * (256 if kind == "surf" else 16) * 1.965e9
    print(f"{kind} pairs={len(pairs)} F={n_feat} sweep_ms={ms:.3f} finalize_ms={s1['last_finalize_ms']:.3f} "
          f"cmp/s={comps / ms * 1e3:.4g} frac={comps * ops / (ms * 1e-3) / peak:.3f} matches={res.n_matches}")
