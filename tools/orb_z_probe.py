"""ORB tensor-core sweep, "Z" operand encoding ($ESFM_ORB_Z=1: packed (distance, column) keys straight from the MMA, branch-free
row selection) against the XOR + POPC engine: byte-identity on a ragged bank (pairs split over several CTAs) and on all pairs of the
timing bank (one CTA per pair: strict column thresholds), then sweep rates of Z off / on.
usage: python tools/orb_z_probe.py [n_images] [n_feat]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import easysfm_b200 as esfm
from easysfm_b200 import scheduler, synth

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
dev = torch.device("cuda:0")
rows = [700, 0, 1, 129, 1025, 2, 512, 300, 2049, 128]
ragged = synth.orb_like(len(rows), rows, seed=9)
ragged[4][10] = ragged[4][3]; ragged[6][5] = ragged[4][3]; ragged[6][7] = ragged[4][3]
ragged[8][2048] = ragged[4][3]; ragged[9][127] = ragged[4][3]
data = synth.orb_like_torch(n_images, n_feat, 5, dev).reshape(-1).view(torch.uint8)
pairs = scheduler.all_pairs(n_images)


def ragged_bytes(ctx):
    bank = ctx.bank_from_frames(ragged)
    out = []
    for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
        res = bank.match_all_pairs(ratio, cc)
        out += [res.pair_at(k)[2].tobytes() for k in range(res.n_pairs)]
        res.close()
    for (i, j) in ((4, 6), (6, 4), (0, 8), (8, 3), (3, 2), (9, 4), (4, 9), (8, 4)):
        idx, dist = bank.knn2_pair(i, j)
        out += [idx.tobytes(), dist.tobytes()]
        out.append(bank.match_pair(i, j, 0.8, True).tobytes())
    bank.close()
    return out


def timing_bank(ctx):
    bank = ctx.bank(esfm.KIND_B256, n_images)
    for f in range(n_images):
        bank.set_frame_rows(f, n_feat)
    bank.alloc_device()
    ptr, nbytes = bank.device_rows()
    scheduler._wrap_device_bytes(ptr, nbytes, 0).copy_(data)
    torch.cuda.synchronize()
    bank.commit_device()
    return bank


def all_pairs_bytes(bank):
    res = bank.match_pairs(pairs, 0.8, True)
    out = [res.pair_at(k)[2].tobytes() for k in range(res.n_pairs)]
    res.close()
    return out


def rate(ctx, bank, label):
    best = 1e30
    for r in range(4):
        res = bank.match_pairs(pairs, 0.8, True, device_resident=True)
        res.close()
        ctx.synchronize()
        if r:
            best = min(best, ctx.stats()["last_sweep_ms"])
    print(f"{label}: {len(pairs)} pairs of {n_feat}x{n_feat}: sweep {best:.2f} ms -> {len(pairs) * n_feat * n_feat / (best * 1e-3):.3e} cmp/s", flush=True)


os.environ.pop("ESFM_ORB_Z", None)
os.environ.pop("ESFM_TC_DEBUG", None)
ref_ctx = esfm.Context(0)
ref_ctx.set_hamming_engine("popc")
ref_ragged = ragged_bytes(ref_ctx)
ref_bank = timing_bank(ref_ctx)
ref_all = all_pairs_bytes(ref_bank)
rate(ref_ctx, ref_bank, "popc      ")
ref_bank.close()
for z in [int(x) for x in os.environ.get("ORB_Z_PROBE_MODES", "0,1").split(",")]:
    os.environ["ESFM_ORB_Z"] = str(z)
    ctx = esfm.Context(0)
    ctx.set_hamming_engine("tc")
    got = ragged_bytes(ctx)
    bad = [k for k, (a, b) in enumerate(zip(got, ref_ragged)) if a != b]
    print(f"z={z}: ragged bank vs popc engine: {'IDENTICAL' if not bad and len(got) == len(ref_ragged) else 'MISMATCH at items ' + str(bad[:10])}", flush=True)
    bank = timing_bank(ctx)
    got_all = all_pairs_bytes(bank)
    badp = [k for k, (a, b) in enumerate(zip(got_all, ref_all)) if a != b]
    print(f"z={z}: all {len(pairs)} pairs of the timing bank vs popc engine: {'IDENTICAL' if not badp else str(len(badp)) + ' pairs differ, first ' + str(badp[:5])}", flush=True)
    if badp:
        k = badp[0]
        a = np.frombuffer(got_all[k], esfm.capi.DMATCH_DTYPE); b = np.frombuffer(ref_all[k], esfm.capi.DMATCH_DTYPE)
        print(f"   pair {pairs[k].tolist()}: {len(a)} vs {len(b)} matches; first rows got {a[:3].tolist()} want {b[:3].tolist()}", flush=True)
        n = min(len(a), len(b))
        d = np.nonzero((a[:n]["queryIdx"] != b[:n]["queryIdx"]) | (a[:n]["trainIdx"] != b[:n]["trainIdx"]) | (a[:n]["distance"] != b[:n]["distance"]))[0]
        if len(d):
            print(f"   first differing row {d[0]}: got {a[d[0]].tolist()} want {b[d[0]].tolist()}", flush=True)
    rate(ctx, bank, f"tc z={z}    ")
    if z == 1 and os.environ.get("ORB_Z_PROBE_DRAIN"):
        for dbg in (1,):
            os.environ["ESFM_TC_DEBUG"] = str(dbg)
            rate(ctx, bank, f"tc z=1 dbg{dbg}")
        os.environ.pop("ESFM_TC_DEBUG", None)
    bank.close()
    ctx.close()
ref_ctx.close()
