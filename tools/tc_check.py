"""Quick GPU check of the tensor-core L2 engine against the FFMA engine and the C oracle (small shapes first)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import easysfm_b200 as esfm
import oracle
from easysfm_b200 import synth

ctx = esfm.Context(0)
shapes = [(128, 128), (1, 2), (37, 53), (257, 1025), (1500, 700), (2049, 3000), (8000, 8000)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
bad = 0
for nq, nt in shapes:
    Q, T = synth.surf_like(2, [nq, nt], seed=nq * 3 + nt)
    out = {}
    for eng in ("ffma", "tc"):
        ctx.set_l2_engine(eng)
        bank = ctx.bank_from_frames([Q, T])
        t0 = time.time()
        idx, dist = bank.knn2_pair(0, 1)
        m = bank.match_pair(0, 1, 0.8, True)
        out[eng] = (idx, dist, m, time.time() - t0)
        bank.close()
    ridx, rdist = oracle.knn2(Q, T)
    for eng in ("ffma", "tc"):
        idx, dist, m, dt = out[eng]
        nidx = int((idx != ridx).sum())
        ok = np.allclose(dist, rdist, rtol=1e-5) if nidx == 0 else False
        print(f"{nq}x{nt} {eng}: idx mismatches vs oracle {nidx}, dist ok {ok}, matches {len(m)}, {dt*1e3:.1f} ms", flush=True)
    a, b = out["ffma"][2], out["tc"][2]
    same = len(a) == len(b) and (a["queryIdx"] == b["queryIdx"]).all() and (a["trainIdx"] == b["trainIdx"]).all()
    print(f"   cross-check matches identical between engines: {same}", flush=True)
    if not same or int((out['tc'][0] != ridx).sum()) > 2:
        bad += 1
print("TC CHECK", "FAILED" if bad else "OK")
sys.exit(1 if bad else 0)
