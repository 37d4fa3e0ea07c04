"""ORB on the tensor cores: geometry and pipeline probes in one process (one torch import, a few seconds of GPU time).
For query blocks of 1 and 2 tiles ($ESFM_TC_QT_ORB, read at esfm_init): byte-identity with the XOR + POPC engine on a ragged
bank, then the sweep rate at the bench frame size with the pipeline probes of sweep_l2_tc.cu ($ESFM_TC_DEBUG, read per launch;
results are WRONG when it is set): 0 = full kernel, 1 = epilogue only drains tensor memory, 5 = additionally no TMA loads.
usage: python tools/orb_tc_probe.py [n_images] [n_feat]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
dev = torch.device("cuda:0")
rows = [700, 0, 1, 129, 1025, 2, 512, 300, 2049]
ragged = synth.orb_like(len(rows), rows, seed=9)
ragged[4][10] = ragged[4][3]; ragged[6][5] = ragged[4][3]; ragged[6][7] = ragged[4][3]
data = synth.orb_like_torch(n_images, n_feat, 5, dev).reshape(-1).view(torch.uint8)
pairs = scheduler.all_pairs(n_images)


def all_bytes(ctx):
    bank = ctx.bank_from_frames(ragged)
    out = []
    for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
        res = bank.match_all_pairs(ratio, cc)
        out += [res.pair_at(k)[2].tobytes() for k in range(res.n_pairs)]
    for (i, j) in ((4, 6), (6, 4), (0, 8), (8, 3), (3, 2)):
        idx, dist = bank.knn2_pair(i, j)
        out += [idx.tobytes(), dist.tobytes()]
    bank.close()
    return out


os.environ.pop("ESFM_TC_DEBUG", None)
ref_ctx = esfm.Context(0)
ref_ctx.set_hamming_engine("popc")
ref = all_bytes(ref_ctx)
ref_counts = None
for qt in (1, 2):
    os.environ["ESFM_TC_QT_ORB"] = str(qt)
    os.environ.pop("ESFM_TC_DEBUG", None)
    ctx = esfm.Context(0)
    ctx.set_hamming_engine("tc")
    same = all_bytes(ctx) == ref
    print(f"qt={qt}: ragged bank vs popc engine: {'IDENTICAL' if same else 'MISMATCH'}", flush=True)
    bank = ctx.bank(esfm.KIND_B256, n_images)
    for f in range(n_images):
        bank.set_frame_rows(f, n_feat)
    bank.alloc_device()
    ptr, nbytes = bank.device_rows()
    scheduler._wrap_device_bytes(ptr, nbytes, 0).copy_(data)
    torch.cuda.synchronize()
    bank.commit_device()
    for dbg in (0, 1, 5):
        if dbg:
            os.environ["ESFM_TC_DEBUG"] = str(dbg)
        else:
            os.environ.pop("ESFM_TC_DEBUG", None)
        best = 1e30
        for r in range(4):
            res = bank.match_pairs(pairs, 0.8, True, device_resident=True)
            if dbg == 0:
                counts = res.pair_counts().copy()
            res.close()
            ctx.synchronize()
            if r:
                best = min(best, ctx.stats()["last_sweep_ms"])
        print(f"qt={qt} debug={dbg}: {len(pairs)} pairs of {n_feat}x{n_feat}: sweep {best:.2f} ms -> "
              f"{len(pairs) * n_feat * n_feat / (best * 1e-3):.3e} cmp/s", flush=True)
    os.environ.pop("ESFM_TC_DEBUG", None)
    if ref_counts is None:
        ref_counts = counts
    else:
        print("qt=2 pair counts equal qt=1 at the timing size:", bool((counts == ref_counts).all()), flush=True)
    bank.close()
    ctx.close()
ref_ctx.close()
