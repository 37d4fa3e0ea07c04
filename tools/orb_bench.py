"""ORB extraction throughput: esfm_orb_extract / esfm_bank_set_frame_from_image on cuda:0 against cv2's ORB (the call the
reference makes, feature_matching.cpp:16-22) on the box's host cores, same seeded images, outputs compared bit for bit on every frame.

    python tools/orb_bench.py [--frames 24] [--out gpurun_out/orb_bench.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from easysfm_b200 import capi  # noqa: E402
from orb_util import image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "orb_bench.json"))
    a = ap.parse_args()
    try:
        import cv2
    except ImportError:
        cv2 = None
    ctx = capi.Context(0)
    report = {"frames": a.frames, "cases": []}
    for name, (h, w, nf) in {"vga_5000": (480, 640, 5000), "hd1080_5000": (1080, 1920, 5000), "hd1080_20000": (1080, 1920, 20000)}.items():
        imgs = [image(900 + k, h, w, 80 if h < 1000 else 400, bgr=True) for k in range(a.frames)]
        ctx.orb_extract(imgs[0], nf)                      # warm-up: scratch allocation
        t0 = time.perf_counter()
        outs = [ctx.orb_extract(im, nf) for im in imgs]
        t_gpu = (time.perf_counter() - t0) / a.frames
        phases = {}
        for im in imgs:
            ctx.orb_extract(im, nf)
            for k, v in ctx.orb_last_timing().items():
                phases[k] = phases.get(k, 0.0) + v / a.frames
        bank = capi.Bank(ctx, capi.KIND_B256, a.frames)
        t0 = time.perf_counter()
        for k, im in enumerate(imgs):
            bank.set_frame_from_image(k, im, nf)
        bank.commit()
        t_bank = (time.perf_counter() - t0) / a.frames
        case = {"case": name, "rows": h, "cols": w, "max_features": nf, "keypoints_mean": float(np.mean([len(k) for k, _ in outs])),
                "gpu_ms_per_frame": t_gpu * 1e3, "gpu_into_bank_ms_per_frame": t_bank * 1e3, "phases_mean": phases, "timing": "host wall clock around the C-ABI call, image in "
                "pageable host memory, key points (and descriptors for the first figure) back on the host"}
        if cv2 is not None:
            det, ext = cv2.ORB_create(nf), cv2.ORB_create(nf)
            t0 = time.perf_counter()
            refs = []
            for im in imgs:
                k = det.detect(im, None)
                refs.append(ext.compute(im, k))
            t_cpu = (time.perf_counter() - t0) / a.frames
            same = 0
            for (kp, d), (rk, rd) in zip(outs, refs):
                ok = len(kp) == len(rk) and np.array_equal(d, rd) and all(
                    (p.pt[0], p.pt[1], p.angle, p.response, p.octave) == (q["x"], q["y"], q["angle"], q["response"], q["octave"]) for p, q in zip(rk, kp))
                same += bool(ok)
            case.update({"cv2_ms_per_frame": t_cpu * 1e3, "cv2_threads": cv2.getNumThreads(), "speedup": t_cpu / t_gpu,
                         "frames_identical_to_cv2": same})
        report["cases"].append(case)
        print(json.dumps(case))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
