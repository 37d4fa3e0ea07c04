"""ORB / Hamming on the tensor cores (FP8 +-1 dot product) against the XOR + POPC engine: every pair of a ragged bank must come
out byte-identical (matches, raw knn-2), then a timing of both engines at the bench frame size.
usage: python tools/ham_tc_check.py [n_images_timing] [n_feat_timing]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

ctx = esfm.Context(0)
rows = [700, 0, 1, 129, 1025, 2, 512, 300, 2049]
frames = synth.orb_like(len(rows), rows, seed=9)
frames[4][10] = frames[4][3]; frames[6][5] = frames[4][3]; frames[6][7] = frames[4][3]      # exact duplicates / zero distances
bad = 0
out = {}
for eng in ("popc", "tc"):
    ctx.set_hamming_engine(eng)
    bank = ctx.bank_from_frames(frames)
    per = {}
    for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
        res = bank.match_all_pairs(ratio, cc)
        per[(ratio, cc)] = [res.pair_at(k) for k in range(res.n_pairs)]
    per["knn"] = [bank.knn2_pair(i, j) for (i, j) in ((4, 6), (6, 4), (0, 8), (8, 3), (3, 2), (0, 5))]
    out[eng] = per
    bank.close()
for key in out["popc"]:
    for a, b in zip(out["popc"][key], out["tc"][key]):
        if key == "knn":
            same = (a[0] == b[0]).all() and (a[1] == b[1]).all()
        else:
            same = a[:2] == b[:2] and a[2].tobytes() == b[2].tobytes()
        if not same:
            bad += 1
            if bad <= 5:
                if key == "knn":
                    d = np.nonzero((a[0] != b[0]).any(axis=1) | (a[1] != b[1]).any(axis=1))[0]
                    print("MISMATCH knn rows", d[:5], "popc", a[0][d[:3]].tolist(), a[1][d[:3]].tolist(), "tc", b[0][d[:3]].tolist(), b[1][d[:3]].tolist())
                else:
                    print("MISMATCH", key, a[:2], "popc n=%d tc n=%d" % (len(a[2]), len(b[2])), a[2][:3], b[2][:3])
print("ham_tc_check:", "IDENTICAL" if bad == 0 else f"{bad} MISMATCHES", flush=True)

import torch
n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
dev = torch.device("cuda:0")
bank = ctx.bank(esfm.KIND_B256, n_images)
for f in range(n_images):
    bank.set_frame_rows(f, n_feat)
bank.alloc_device()
ptr, nbytes = bank.device_rows()
raw = scheduler._wrap_device_bytes(ptr, nbytes, 0)
raw.copy_(synth.orb_like_torch(n_images, n_feat, 5, dev).reshape(-1).view(torch.uint8))
torch.cuda.synchronize()
bank.commit_device()
pairs = scheduler.all_pairs(n_images)
counts = {}
for eng in ("popc", "tc"):
    ctx.set_hamming_engine(eng)
    for r in range(3):
        res = bank.match_pairs(pairs, 0.8, True, device_resident=True)
        counts[eng] = res.pair_counts().copy()
        res.close()
        ctx.synchronize()
        st = ctx.stats()
    print(f"orb engine={eng} {len(pairs)} pairs of {n_feat}x{n_feat}: finalize {st['last_finalize_ms']:.2f} ms, sweep {st['last_sweep_ms']:.2f} ms -> "
          f"{len(pairs) * n_feat * n_feat / (st['last_sweep_ms'] * 1e-3):.3e} cmp/s", flush=True)
print("timing-run pair counts identical:", bool((counts["popc"] == counts["tc"]).all()))
