"""CPU restatement (numpy) of the ORB stage that feeds the matching hot path -- SURVEY.md section 8(f) rank 4.

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and tools/ may import this file; the product path
(easysfm_b200/csrc/orb.cu behind esfm_orb_*) never does.

What the reference does: cpp_code/src/feature_matching.cpp:14-41 (FeatureMatching::detectFeaturesORB) builds
cv::ORB::create(max_num) twice and calls detector->detect(image, keypoints) and descriptor->compute(image, keypoints,
descriptors); cpp_code/test/sfm.cpp:116 calls it once per frame.  The algorithm itself lives in OpenCV (features2d: orb.cpp,
fast.cpp, fast_score.cpp; imgproc: resize.cpp INTER_LINEAR_EXACT, smooth GaussianBlur bit-exact u8 path, color BGR2GRAY; core:
fastAtan2, KeyPointsFilter), a dependency that is not under /root/reference.  This file restates OpenCV 4.13.0's published
algorithm with the defaults of ORB::create(nfeatures): scaleFactor 1.2, 8 levels, edgeThreshold 31, firstLevel 0, WTA_K 2,
HARRIS_SCORE, patchSize 31, fastThreshold 20.

Pinned: tests/golden/make_golden_orb.py runs cv2 4.13.0 in the build container on seeded images and commits key points and
descriptors (tests/golden/orb_extract.npz); tests/test_orb_oracle.py checks this file against them bit for bit (and against
cv2 itself when it is importable).  The 256 sampling pairs were measured from cv2 (tools/probe_orb_pattern.py).
"""
import os

import numpy as np

N_LEVELS = 8
SCALE_FACTOR = float(np.float32(1.2))     # ORB::create takes a float 1.2f and keeps it as a double
EDGE_THRESHOLD = 31
PATCH_SIZE = 31
HALF_PATCH = 15
FAST_THRESHOLD = 20
HARRIS_BLOCK = 7
HARRIS_K = np.float32(0.04)

_PATTERN = None


def pattern():
    """256 x (x0, y0, x1, y1), int32 -- orb.cpp bit_pattern_31_ as measured by tools/probe_orb_pattern.py."""
    global _PATTERN
    if _PATTERN is None:
        _PATTERN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_pattern.npy")).astype(np.int32)
    return _PATTERN


def bgr_to_gray(bgr):
    """imgproc color_rgb: 15-bit fixed point, B 3735, G 19235, R 9798 (ORB converts a 3-channel input before anything else)."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _linear_coeffs(src, dst):
    scale = src / dst
    d = np.arange(dst, dtype=np.float64)
    fx = (d + 0.5) * scale - 0.5
    sx = np.floor(fx)
    f = fx - sx
    lo = sx < 0
    hi = sx >= src - 1
    sx = np.where(lo, 0, np.where(hi, src - 1, sx)).astype(np.int64)
    f = np.where(lo | hi, 0.0, f)
    return sx, np.rint(f * 256).astype(np.int64)


def resize_linear_exact(img, dw, dh):
    """resize.cpp, INTER_LINEAR_EXACT on u8: 8.8 fixed-point weights per axis, horizontal pass exact in 8.8, vertical pass in
    16.16, one rounding at the end."""
    sh, sw = img.shape
    ix, cx = _linear_coeffs(sw, dw)
    iy, cy = _linear_coeffs(sh, dh)
    a = img.astype(np.int64)
    ix1 = np.minimum(ix + 1, sw - 1)
    iy1 = np.minimum(iy + 1, sh - 1)
    h = a[:, ix] * (256 - cx) + a[:, ix1] * cx
    v = h[iy, :] * (256 - cy)[:, None] + h[iy1, :] * cy[:, None]
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def gaussian_kernel_7():
    """getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) in double, normalised, then cast."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(np.float32)


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in the 64-bit mantissa of long double, so one rounding."""
    ld = np.longdouble
    return (np.asarray(a, np.float32).astype(ld) * np.asarray(b, np.float32).astype(ld) + np.asarray(c, np.float32).astype(ld)).astype(np.float32)


def gaussian_blur_7x7(img):
    """GaussianBlur(Size(7, 7), 2, 2, BORDER_REFLECT_101) as ORB calls it -- on a SUB-MATRIX of its pyramid buffer, which takes
    smooth.dispatch.cpp past the bit-exact fixed-point branch (that one requires !isSubmatrix()) into sepFilter2D with the float32
    kernel (filter.simd.hpp RowFilter<uchar, float>, SymmColumnFilter<Cast<float, uchar>>): row pass s = k[0] p[0], then
    s = fma(k[i], p[i], s) left to right; column pass s = k[3] r[0], then s = fma(k[3 + j], r[+j] + r[-j], s); round-half-even and
    saturate.  The fused multiply-adds are what the AVX2 / AVX-512 dispatch of that file executes (-mfma contracts `s += f * S[i]`);
    cv2 on an SSE-only CPU would run the baseline copy with separate roundings and differ on exact .5 ties (1 bit in 3 million on
    the golden images).  cv2.GaussianBlur on a whole image takes the fixed-point branch instead and differs from this by one grey
    level in places; the descriptors tell the three apart, and only this one reproduces tests/golden/orb_extract.npz."""
    k = gaussian_kernel_7()
    a = np.pad(img.astype(np.float32), 3, mode="reflect")
    h, w = img.shape
    rows = k[0] * a[:, 0:w]
    for i in range(1, 7):
        rows = _fma32(k[i], a[:, i:i + w], rows)
    v = k[3] * rows[3:3 + h]
    for j in range(1, 4):
        v = _fma32(k[3 + j], rows[3 + j:3 + j + h] + rows[3 - j:3 - j + h], v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def level_scales(n_levels=N_LEVELS):
    """orb.cpp getScale: (float)pow(scaleFactor, level - firstLevel)."""
    return [np.float32(SCALE_FACTOR ** l) for l in range(n_levels)]


def c_round(x):
    """cvRound: round half to even (cvtsd2si)."""
    return int(np.rint(np.float64(x)))


def build_pyramid(gray, n_levels=N_LEVELS):
    """orb.cpp detectAndCompute: level 0 is the image; level l is resized from level l - 1 to
    (cvRound(cols * (1.f / scale_l)), cvRound(rows * (1.f / scale_l)))."""
    scales = level_scales(n_levels)
    h, w = gray.shape
    levels = [gray]
    for l in range(1, n_levels):
        inv = np.float32(1.0) / scales[l]
        dw = c_round(np.float32(w) * inv)          # cols * inv_scale in float: 297 rows at level 1 give 247, not 247.5 -> 248
        dh = c_round(np.float32(h) * inv)
        levels.append(resize_linear_exact(levels[-1], dw, dh))
    return levels, scales


def compute_descriptors(levels, scales, kps):
    """orb.cpp computeOrbDescriptors on the blurred levels.  kps: structured array with x, y (full-resolution coordinates), octave,
    angle (degrees).  Returns uint8 [n, 32]."""
    pat = pattern()
    blurred = {}
    out = np.zeros((len(kps), 32), np.uint8)
    px0 = pat[:, 0].astype(np.float32); py0 = pat[:, 1].astype(np.float32)
    px1 = pat[:, 2].astype(np.float32); py1 = pat[:, 3].astype(np.float32)
    for i, kp in enumerate(kps):
        l = int(kp["octave"])
        if l not in blurred:
            blurred[l] = gaussian_blur_7x7(levels[l])
        img = blurred[l]
        inv = np.float32(1.0) / scales[l]
        cx = c_round(np.float32(kp["x"]) * inv)
        cy = c_round(np.float32(kp["y"]) * inv)
        ang = np.float32(kp["angle"]) * np.float32(np.pi / 180.0)
        a = np.float32(np.cos(np.float64(ang)))
        b = np.float32(np.sin(np.float64(ang)))

        def sample(px, py):
            x = px * a - py * b                      # float32, two roundings per product and one for the difference
            y = px * b + py * a
            ix = np.rint(x).astype(np.int64)
            iy = np.rint(y).astype(np.int64)
            return img[cy + iy, cx + ix]
        t0 = sample(px0, py0)
        t1 = sample(px1, py1)
        out[i] = np.packbits((t0 < t1).astype(np.uint8), bitorder="little")
    return out


# ---- detection (orb.cpp computeKeyPoints) ----------------------------------------------------------------------------------------

# fast.cpp makeOffsets, pattern size 16: the Bresenham circle of radius 3, clockwise from (0, 3)
FAST_CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
               (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_scores(img, threshold=FAST_THRESHOLD):
    """fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>: int32 map, 0 where the pixel is no corner, else the largest threshold
    the pixel would still pass (max over the 16 nine-pixel arcs of the arc's weakest contrast, either polarity, minus 1).  Rows and
    columns 0..2 and the last three are never tested."""
    h, w = img.shape
    a = img.astype(np.int32)
    v = a[3:h - 3, 3:w - 3]
    d = np.stack([v - a[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for (dx, dy) in FAST_CIRCLE])     # [16, h-6, w-6]
    dd = np.concatenate([d, d[:8]])                                                              # arcs wrap
    best_a = np.full(v.shape, -(1 << 20), np.int32)
    best_b = np.full(v.shape, -(1 << 20), np.int32)
    for k in range(16):
        arc = dd[k:k + 9]
        best_a = np.maximum(best_a, arc.min(0))          # centre brighter than the whole arc
        best_b = np.maximum(best_b, (-arc).min(0))       # centre darker
    s = np.maximum(best_a, best_b)
    out = np.zeros((h, w), np.int32)
    out[3:h - 3, 3:w - 3] = np.where(s > threshold, s - 1, 0)
    return out


def fast_nms(score):
    """fast.cpp: a corner survives when its score is strictly greater than its eight neighbours' (0 for non-corners).  Returns
    (y, x) in raster order, which is the order FAST emits them in."""
    h, w = score.shape
    c = score[1:h - 1, 1:w - 1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy or dx:
                keep &= c > score[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    ys, xs = np.nonzero(keep)
    return ys + 1, xs + 1


def features_per_level(n_features, n_levels=N_LEVELS):
    """orb.cpp computeKeyPoints: geometric split of nfeatures over the levels, float arithmetic as written there."""
    factor = np.float32(1.0 / SCALE_FACTOR)
    per = np.float32(n_features) * (np.float32(1) - factor) / (np.float32(1) - np.float32(np.float64(factor) ** np.float64(n_levels)))
    out, total = [], 0
    for _ in range(n_levels - 1):
        n = c_round(per)
        out.append(n)
        total += n
        per = per * factor
    out.append(max(n_features - total, 0))
    return out


_SELECT = None


def retain_best(responses, n_points):
    """KeyPointsFilter::retainBest (oracle/orb_select.cpp): surviving indices in the order the filter leaves them."""
    global _SELECT
    import ctypes
    if _SELECT is None:
        here = os.path.dirname(os.path.abspath(__file__))
        if not os.path.exists(os.path.join(here, "liborbselect.so")):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", here, "liborbselect.so"])
        _SELECT = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "liborbselect.so"))
        _SELECT.orb_oracle_retain_best.restype = ctypes.c_int
    r = np.ascontiguousarray(responses, np.float32)
    out = np.zeros(max(len(r), 1), np.int32)
    n = _SELECT.orb_oracle_retain_best(r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(r)), ctypes.c_int(int(n_points)),
                                       out.ctypes.data_as(ctypes.c_void_p))
    return out[:n].copy()


def harris_responses(img, xs, ys):
    """orb.cpp HarrisResponses, blockSize 7, k 0.04: integer Sobel-like gradients summed over the 7 x 7 block, one float expression."""
    a = img.astype(np.int64)
    r = HARRIS_BLOCK // 2
    scale = np.float32(1.0) / (np.float32(4 * HARRIS_BLOCK) * np.float32(255.0))
    scale4 = scale * scale * scale * scale
    out = np.zeros(len(xs), np.float32)
    for i, (x0, y0) in enumerate(zip(xs, ys)):
        p = a[y0 - r - 1:y0 + r + 2, x0 - r - 1:x0 + r + 2]         # 9 x 9
        ix = (p[1:-1, 2:] - p[1:-1, :-2]) * 2 + (p[:-2, 2:] - p[:-2, :-2]) + (p[2:, 2:] - p[2:, :-2])
        iy = (p[2:, 1:-1] - p[:-2, 1:-1]) * 2 + (p[2:, :-2] - p[:-2, :-2]) + (p[2:, 2:] - p[:-2, 2:])
        sa = int((ix * ix).sum()); sb = int((iy * iy).sum()); sc = int((ix * iy).sum())
        fa, fb, fc = np.float32(sa), np.float32(sb), np.float32(sc)
        out[i] = (fa * fb - fc * fc - HARRIS_K * (fa + fb) * (fa + fb)) * scale4
    return out


def _umax():
    """orb.cpp: half-widths of the circular patch rows (radius 15), made symmetric."""
    half = HALF_PATCH
    umax = [0] * (half + 2)
    vmax = int(np.floor(half * np.sqrt(np.float32(2.0)) / 2 + 1))
    vmin = int(np.ceil(half * np.sqrt(np.float32(2.0)) / 2))
    for v in range(vmax + 1):
        umax[v] = c_round(np.sqrt(float(half * half - v * v)))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax[:half + 1]


UMAX = _umax()


def fast_atan2(y, x):
    """core mathfuncs_core: the degree-valued polynomial arctangent (float32 throughout)."""
    f = np.float32
    scale = f(180.0 / np.pi)
    p1 = f(0.9997878412794807) * scale
    p3 = f(-0.3258083974640975) * scale
    p5 = f(0.1555786518463281) * scale
    p7 = f(-0.04432655554792128) * scale
    eps = f(2.220446049250313e-16)
    x = f(x); y = f(y)
    ax, ay = abs(x), abs(y)
    if ax >= ay:
        c = ay / (ax + eps)
        c2 = c * c
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    else:
        c = ax / (ay + eps)
        c2 = c * c
        a = f(90.0) - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    if x < 0:
        a = f(180.0) - a
    if y < 0:
        a = f(360.0) - a
    return f(a)


def ic_angles(img, xs, ys):
    """orb.cpp ICAngles: intensity-centroid orientation over the circular patch of radius 15 on the un-blurred level."""
    a = img.astype(np.int64)
    out = np.zeros(len(xs), np.float32)
    us = {d: np.arange(-d, d + 1) for d in set(UMAX)}
    for i, (x0, y0) in enumerate(zip(xs, ys)):
        u = us[HALF_PATCH]
        m10 = int((u * a[y0, x0 - HALF_PATCH:x0 + HALF_PATCH + 1]).sum())
        m01 = 0
        for v in range(1, HALF_PATCH + 1):
            d = UMAX[v]
            plus = a[y0 + v, x0 - d:x0 + d + 1]
            minus = a[y0 - v, x0 - d:x0 + d + 1]
            m01 += v * int((plus - minus).sum())
            m10 += int((us[d] * (plus + minus)).sum())
        out[i] = fast_atan2(np.float32(m01), np.float32(m10))
    return out


KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])


def detect(gray, n_features, n_levels=N_LEVELS):
    """cv::ORB::detect with the reference's parameters (feature_matching.cpp:16,20).  Returns (key points, levels, scales); the key
    points come out level by level, inside a level in the order KeyPointsFilter::retainBest leaves them."""
    levels, scales = build_pyramid(gray, n_levels)
    per_level = features_per_level(n_features, n_levels)
    chunks = []
    for l, img in enumerate(levels):
        h, w = img.shape
        if h <= 2 * EDGE_THRESHOLD or w <= 2 * EDGE_THRESHOLD or h < 7 or w < 7:
            continue
        score = fast_scores(img)
        ys, xs = fast_nms(score)
        inside = (xs >= EDGE_THRESHOLD) & (xs < w - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < h - EDGE_THRESHOLD)
        ys, xs = ys[inside], xs[inside]
        resp = score[ys, xs].astype(np.float32)
        keep = retain_best(resp, 2 * per_level[l])
        xs, ys = xs[keep], ys[keep]
        harris = harris_responses(img, xs, ys)
        keep = retain_best(harris, per_level[l])
        xs, ys, harris = xs[keep], ys[keep], harris[keep]
        ang = ic_angles(img, xs, ys)
        k = np.zeros(len(xs), KP_DTYPE)
        k["x"] = xs.astype(np.float32) * scales[l]
        k["y"] = ys.astype(np.float32) * scales[l]
        k["size"] = np.float32(PATCH_SIZE) * scales[l]
        k["angle"] = ang
        k["response"] = harris
        k["octave"] = l
        chunks.append(k)
    kps = np.concatenate(chunks) if chunks else np.zeros(0, KP_DTYPE)
    return kps, levels, scales


def detect_and_compute(image, n_features):
    """FeatureMatching::detectFeaturesORB (feature_matching.cpp:14-41): detect, then compute on the same image (BGR or gray)."""
    gray = bgr_to_gray(image) if image.ndim == 3 else image
    kps, levels, scales = detect(gray, n_features)
    return kps, compute_descriptors(levels, scales, kps)
