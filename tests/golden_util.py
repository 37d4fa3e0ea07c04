import os

import numpy as np

from oracle import DMATCH_DTYPE

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def as_matches(idx, dist):
    out = np.zeros(len(idx), DMATCH_DTYPE)
    if len(idx):
        out["queryIdx"] = idx[:, 0]
        out["trainIdx"] = idx[:, 1]
        out["distance"] = dist
    return out


def frames_of(z):
    n = len([k for k in z.files if k.startswith("frame_")])
    return [z[f"frame_{i}"] for i in range(n)]
