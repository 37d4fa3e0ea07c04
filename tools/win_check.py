"""Quick GPU check of the 16-bit tensor-core sweeps (engine tc16) against the oracle and the other engines, then their rates at the
bench shapes.  usage: python tools/win_check.py [orb|surf|both]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import easysfm_b200 as esfm
import oracle
from easysfm_b200 import synth

which = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = esfm.Context(0)
bad = 0
for kind in ("orb", "surf"):
    if which not in (kind, "both"):
        continue
    gen = synth.orb_like if kind == "orb" else synth.surf_like
    setter = ctx.set_hamming_engine if kind == "orb" else ctx.set_l2_engine
    for shape in ([128, 128], [1, 2], [37, 53], [257, 1025], [1500, 700], [2049, 3000]):
        frames = gen(2, shape, seed=3)
        for (ratio, cc) in ((0.8, True), (0.8, False), (float("inf"), True)):
            ref = oracle.match(frames[0], frames[1], ratio, cc)
            setter("tc16")
            got = ctx.match_descriptors(frames[0], frames[1], ratio, cc)
            same = len(got) == len(ref) and (got["queryIdx"] == ref["queryIdx"]).all() and (got["trainIdx"] == ref["trainIdx"]).all()
            if same and kind == "orb":
                same = (got["distance"] == ref["distance"]).all()
            elif same:
                same = np.allclose(got["distance"], ref["distance"], rtol=1e-5)
            if not same:
                bad += 1
                q_ref = dict(zip(ref["queryIdx"].tolist(), zip(ref["trainIdx"].tolist(), ref["distance"].tolist())))
                q_got = dict(zip(got["queryIdx"].tolist(), zip(got["trainIdx"].tolist(), got["distance"].tolist())))
                diff = [(q, q_ref.get(q), q_got.get(q)) for q in sorted(set(q_ref) | set(q_got)) if q_ref.get(q) != q_got.get(q)]
                print(f"{kind} {shape} ratio={ratio} cc={cc}: MISMATCH ref {len(ref)} got {len(got)}; first diffs (q, ref, got): {diff[:6]}", flush=True)
            else:
                print(f"{kind} {shape} ratio={ratio} cc={cc}: ok ({len(ref)} matches)", flush=True)
    # all pairs of a small ragged bank, one CTA per pair path too (enough pairs) + knn2
    rows = [700, 0, 1, 129, 1025, 2, 512, 300]
    frames = gen(len(rows), rows, seed=9)
    out = {}
    for eng in (("popc", "tc", "tc16") if kind == "orb" else ("ffma", "tc", "tc16")):
        setter(eng)
        bank = ctx.bank_from_frames(frames)
        per = []
        for ratio, cc in ((0.8, True), (0.8, False), (float("inf"), True)):
            res = bank.match_all_pairs(ratio, cc)
            per += [res.pair_at(k)[2] for k in range(res.n_pairs)]
        for (i, j) in ((4, 6), (6, 4), (0, 7), (3, 2), (0, 5), (2, 0)):
            idx, dist = bank.knn2_pair(i, j)
            per += [idx, dist]
        out[eng] = per
        bank.close()
    base = "popc" if kind == "orb" else "ffma"
    nd = sum(1 for a, b in zip(out[base], out["tc16"]) if a.tobytes() != b.tobytes())
    print(f"{kind}: ragged bank, {len(out[base])} result arrays, tc16 differs from {base} in {nd}", flush=True)
    if nd:
        bad += 1
        for n, (a, b) in enumerate(zip(out[base], out["tc16"])):
            if a.tobytes() != b.tobytes():
                print("   first differing array", n, a.dtype, a.shape, b.shape, a[:4], b[:4])
                break
print("WIN CHECK", "FAILED" if bad else "OK", flush=True)
ctx.close()
