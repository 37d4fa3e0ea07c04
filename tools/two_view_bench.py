"""Throughput of esfm_two_view_batch next to the reference's cv2 calls on the host: P synthetic two-view scenes of M matches
(pixel noise 0.5, 30 % gross outliers).  usage: python tools/two_view_bench.py [pairs] [matches]  -> one JSON line"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import easysfm_b200 as esfm
from two_view_util import scene, rot_angle_deg, dir_angle_deg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
base = [scene(M, 0.5, 0.3, 900 + k) for k in range(16)]            # 16 distinct scenes, repeated (the sampler key differs per pair)
K = base[0][0]
p1 = np.concatenate([base[k % 16][1] for k in range(P)])
p2 = np.concatenate([base[k % 16][2] for k in range(P)])
off = (np.arange(P + 1) * M).astype(np.int64)
ctx = esfm.Context(0)
ctx.two_view_batch(off[:65], p1[:64 * M], p2[:64 * M], K)           # warm-up
t0 = time.perf_counter()
mask, res = ctx.two_view_batch(off, p1, p2, K, 1.0, 0.99, max_iters=1000, seed=1)
dt = time.perf_counter() - t0
rerr = [rot_angle_deg(res[k]["R"], base[k % 16][3]) for k in range(P)]
terr = [dir_angle_deg(res[k]["t"], base[k % 16][4]) for k in range(P)]
out = {"pairs": P, "matches_per_pair": M, "seconds": dt, "pairs_per_s": P / dt, "hypotheses_mean": float(res["iters"].mean()),
       "inliers_mean": float(res["n_inliers"].mean()), "rot_err_deg_median": float(np.median(rerr)), "t_err_deg_median": float(np.median(terr)),
       "ok": int(res["ok"].sum()), "note": "host arrays in, inlier masks + poses out (h2d/d2h inside the time)"}
try:
    import cv2
    n = min(P, 48)
    t0 = time.perf_counter()
    rr, tt = [], []
    for k in range(n):
        _, x1, x2, R, t = base[k % 16]
        E, m = cv2.findEssentialMat(x1, x2, K, cv2.RANSAC, 0.99, 1.0)
        _, Rc, tc, _ = cv2.recoverPose(E, x1, x2, K, mask=m)
        rr.append(rot_angle_deg(Rc, R)); tt.append(dir_angle_deg(tc, t))
    dc = time.perf_counter() - t0
    out["cpu_reference"] = {"pairs": n, "seconds": dc, "pairs_per_s": n / dc, "threads": cv2.getNumThreads(), "rot_err_deg_median": float(np.median(rr)),
                            "t_err_deg_median": float(np.median(tt)), "what": "cv2.findEssentialMat(RANSAC, 0.99, 1.0) + cv2.recoverPose per pair"}
except ImportError:
    pass
print(json.dumps(out))
