"""SURF tensor-core sweep: generic selection epilogue vs the EXPERIMENTAL branch-free row selection ($ESFM_TC_SURF_BF=1, read at
esfm_init; written in round 1 without GPU time left to run it).  For both: knn-2 against the C oracle on several shapes (index
mismatches must be float64 near-ties), exact-tie data (quantised descriptors, duplicates: indices must be identical), matches with
ratio + cross-check against the FFMA engine, then the sweep rate at the bench frame size.
usage: python tools/surf_bf_probe.py [n_images] [n_feat]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
import oracle
from easysfm_b200 import scheduler, synth

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 38
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
dev = torch.device("cuda:0")
shapes = [(128, 128), (1, 2), (37, 53), (257, 1025), (1500, 700), (2049, 3000)]
data = synth.surf_like_torch(n_images, n_feat, 4, dev).reshape(-1).view(torch.uint8)
pairs = scheduler.all_pairs(n_images)


def quantised(nq, nt, levels):
    rng = np.random.default_rng(nq * 31 + nt)
    Q = (rng.integers(-levels, levels + 1, (nq, 64)) / 8.0).astype(np.float32)
    T = (rng.integers(-levels, levels + 1, (nt, 64)) / 8.0).astype(np.float32)
    T[7] = T[3]; T[nt - 1] = T[3]; Q[11] = T[3]; Q[12] = T[3]
    return Q, T


ref_ctx = esfm.Context(0)
ref_ctx.set_l2_engine("ffma")
for bf in (0, 1):
    os.environ["ESFM_TC_SURF_BF"] = str(bf)
    ctx = esfm.Context(0)
    ctx.set_l2_engine("tc")
    bad = 0
    for nq, nt in shapes:
        Q, T = synth.surf_like(2, [nq, nt], seed=nq * 3 + nt)
        bank = ctx.bank_from_frames([Q, T])
        idx, dist = bank.knn2_pair(0, 1)
        m = bank.match_pair(0, 1, 0.8, True)
        bank.close()
        ridx, rdist = oracle.knn2(Q, T)
        rows = np.nonzero((idx != ridx).any(axis=1))[0]
        # an index mismatch is acceptable only where the float64 distances of the two candidates agree to 1e-5 relative
        D = ((Q[rows, None, :].astype(np.float64) - T[None, :, :]) ** 2).sum(-1) ** 0.5 if len(rows) else None
        unjust = 0
        for k, r in enumerate(rows):
            for c in range(min(2, nt)):
                if idx[r, c] != ridx[r, c] and abs(D[k, idx[r, c]] - D[k, ridx[r, c]]) > 1e-5 * D[k, ridx[r, c]]:
                    unjust += 1
        mr = ref_ctx.match_descriptors(Q, T, 0.8, True)
        same = len(m) == len(mr) and (m["queryIdx"] == mr["queryIdx"]).all() and (m["trainIdx"] == mr["trainIdx"]).all()
        print(f"bf={bf} {nq}x{nt}: rows with an index mismatch vs oracle {len(rows)} (unjustified {unjust}); matches identical to ffma engine: {same}", flush=True)
        bad += unjust + (0 if same or len(rows) else 1)
    for (nq, nt, lv) in ((300, 700, 3), (129, 1025, 2), (515, 260, 5)):
        Q, T = quantised(nq, nt, lv)
        bank = ctx.bank_from_frames([Q, T])
        idx, dist = bank.knn2_pair(0, 1)
        bank.close()
        ridx, rdist = oracle.knn2(Q, T)
        ok = (idx == ridx).all() and (dist == rdist).all()
        okm = all(ctx.match_descriptors(Q, T, r, c).tobytes() == oracle.match(Q, T, r, c).tobytes() for r in (0.8, 1.0, float("inf")) for c in (False, True))
        print(f"bf={bf} exact ties {nq}x{nt}: knn2 identical {bool(ok)}, matches identical {okm}", flush=True)
        bad += (not ok) + (not okm)
    bank = ctx.bank(esfm.KIND_F32X64, n_images)
    for f in range(n_images):
        bank.set_frame_rows(f, n_feat)
    bank.alloc_device()
    ptr, nbytes = bank.device_rows()
    scheduler._wrap_device_bytes(ptr, nbytes, 0).copy_(data)
    torch.cuda.synchronize()
    bank.commit_device()
    best, counts = 1e30, None
    for r in range(4):
        res = bank.match_pairs(pairs, 0.8, True, device_resident=True)
        counts = res.pair_counts().copy()
        res.close()
        ctx.synchronize()
        if r:
            best = min(best, ctx.stats()["last_sweep_ms"])
    print(f"bf={bf}: {len(pairs)} pairs of {n_feat}x{n_feat}: sweep {best:.2f} ms -> {len(pairs) * n_feat * n_feat / (best * 1e-3):.3e} cmp/s; "
          f"matches {int(counts.sum())}; {'PARITY OK' if not bad else str(bad) + ' PARITY PROBLEMS'}", flush=True)
    bank.close()
    ctx.close()
ref_ctx.close()
