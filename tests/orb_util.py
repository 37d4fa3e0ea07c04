"""Seeded synthetic images for the ORB extraction tests (no datasets are reachable): smooth noise fields with rectangles, discs and line
segments drawn on top, so that FAST finds corners at every pyramid level and many of them tie on the integer FAST score."""
import numpy as np


def _smooth(rng, h, w, passes):
    a = rng.random((h, w)).astype(np.float64)
    for _ in range(passes):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    a -= a.min()
    return a / max(a.max(), 1e-9)


def image(seed, h=480, w=640, shapes=60, bgr=False):
    rng = np.random.default_rng(seed)
    img = _smooth(rng, h, w, 6) * 255.0
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(shapes):
        kind = rng.integers(0, 3)
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        amp = float(rng.integers(-90, 90))
        if kind == 0:
            ww, hh = rng.integers(6, 70, 2)
            img[y:y + hh, x:x + ww] += amp
        elif kind == 1:
            r = int(rng.integers(4, 30))
            img[(xx - x) ** 2 + (yy - y) ** 2 <= r * r] += amp
        else:
            th = rng.random() * np.pi
            d = np.abs((xx - x) * np.sin(th) - (yy - y) * np.cos(th))
            img[(d < 1.5) & (np.abs(xx - x) < 60) & (np.abs(yy - y) < 60)] += amp
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if not bgr:
        return g
    out = np.stack([np.clip(g.astype(np.int32) + rng.integers(-20, 20, g.shape), 0, 255) for _ in range(3)], -1)
    return out.astype(np.uint8)


# name -> (seed, h, w, shapes, bgr, nfeatures): the cases of tests/golden/orb_extract.npz
CASES = {
    "vga_5000": (11, 480, 640, 60, False, 5000),      # the reference's default budget (feature_matching.h:13)
    "vga_500": (12, 480, 640, 60, False, 500),
    "odd_bgr_2000": (13, 397, 531, 50, True, 2000),   # odd sizes, 3-channel input (sfm.cpp reads colour images)
    "small_300": (14, 120, 160, 12, False, 300),      # the top levels fall under 2 x edgeThreshold and are skipped
    "sparse_4000": (15, 300, 400, 4, False, 4000),    # fewer corners than the budget: retainBest is a no-op on most levels
    "tiny_100": (16, 70, 90, 3, False, 100),
    "tie_size_3000": (17, 297, 700, 45, False, 3000),  # 297 * (1 / 1.2f) = 247.49999: the level sizes come from a float product, not a quotient
}
