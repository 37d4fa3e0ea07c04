"""Diagnostic: exact-tie L2 case (tests/test_gpu_parity.py::test_l2_exact_ties_on_quantised_descriptors) repeated on both engines,
single-pair launches (pair split over several CTAs) and a 600-copy batch (one CTA per pair); prints every deviation from the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import easysfm_b200 as esfm
import oracle

def data(nq, nt, levels):
    rng = np.random.default_rng(nq * 31 + nt)
    Q = (rng.integers(-levels, levels + 1, (nq, 64)) / 8.0).astype(np.float32)
    T = (rng.integers(-levels, levels + 1, (nt, 64)) / 8.0).astype(np.float32)
    T[7] = T[3]; T[nt - 1] = T[3]; Q[11] = T[3]; Q[12] = T[3]
    return Q, T

def diff(m, ref):
    a = {(int(x["queryIdx"]), int(x["trainIdx"])) for x in m}
    b = {(int(x["queryIdx"]), int(x["trainIdx"])) for x in ref}
    return sorted(a - b), sorted(b - a)

ctx = esfm.Context(0)
for (nq, nt, lv) in ((300, 700, 3), (129, 1025, 2), (515, 260, 5)):
    Q, T = data(nq, nt, lv)
    D = ((Q[:, None, :].astype(np.float64) - T[None, :, :]) ** 2).sum(-1)
    refs = {(r, c): oracle.match(Q, T, r, c) for r in (1.0, float("inf")) for c in (False, True) if not (r == float("inf") and not c)}
    ridx, rdist = oracle.knn2(Q, T)
    for eng in ("ffma", "tc"):
        ctx.set_l2_engine(eng)
        bank = ctx.bank_from_frames([Q, T])
        nbad = 0
        for rep in range(8):
            for (r, c), ref in refs.items():
                for how in ("desc", "bank"):
                    m = ctx.match_descriptors(Q, T, r, c) if how == "desc" else bank.match_pair(0, 1, r, c)
                    extra, missing = diff(m, ref)
                    if extra or missing:
                        nbad += 1
                        if nbad <= 6:
                            print(f"{nq}x{nt} {eng} rep{rep} ratio={r} cc={c} {how}: extra={extra} missing={missing}", flush=True)
                            for (q, t) in extra + missing:
                                tied = np.nonzero(D[:, t] == D[:, t].min())[0]
                                print(f"    q={q} t={t} d2={D[q, t]} colmin={D[:, t].min()} tied queries={tied.tolist()} oracle knn2 of q={ridx[q].tolist()}", flush=True)
            idx, dist = bank.knn2_pair(0, 1)
            if (idx != ridx).any() or (dist != rdist).any():
                nbad += 1
                print(f"{nq}x{nt} {eng} rep{rep} knn2 differs at rows {np.nonzero((idx != ridx).any(axis=1))[0][:8].tolist()}", flush=True)
        # one CTA per pair: 600 copies of the same pair in one batch
        res = bank.match_pairs([[0, 1]] * 600, float("inf"), True)
        ref = refs[(float("inf"), True)]
        nb = 0
        for k in range(600):
            m = res.pair_at(k)[2]
            if len(m) != len(ref) or (m["trainIdx"] != ref["trainIdx"]).any() or (m["queryIdx"] != ref["queryIdx"]).any():
                nb += 1
                if nb <= 2:
                    print(f"{nq}x{nt} {eng} batch copy {k}: {diff(m, ref)}", flush=True)
        print(f"{nq}x{nt} {eng}: {nbad} deviating single-pair calls of {8 * (2 * len(refs) + 1)}, {nb} deviating copies of 600 in the batch", flush=True)
        res.close()
        bank.close()
ctx.close()
