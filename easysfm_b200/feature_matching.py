"""Host-side mirror of the reference's matching interface, running on the CUDA library.

Mirrors, name for name:
  p3dv::FeatureMatching::matchFeaturesORB / matchFeaturesSURF
      cpp_code/include/feature_matching.h:17-21, cpp_code/src/feature_matching.cpp:71-158
      (bool return, matches APPENDED to the caller's list, ratio defaults 0.8 / 0.5, `show` accepted and ignored:
       it only opens GUI windows in the reference, :99-110)
  p3dv::FeatureMatching::detectFeaturesORB
      cpp_code/include/feature_matching.h:13, cpp_code/src/feature_matching.cpp:14-41
      (cv::ORB::create(max_num)->detect + ->compute; fills frame.keypoints and frame.descriptors; `show` ignored)
  frame_t (only the fields the path touches)   cpp_code/include/utility.h:21-54
  pairwise_match's two matcher strategies      python_code/feature_match.py:24-39
      ('mutual_nn' = BFMatcher(crossCheck=True).match sorted by distance, 'ratio_test' = knn-2 + ratio 0.7)

Differences, all deliberate:
  * matchFeaturesSURF is an exact L2 brute-force search, not the approximate FLANN call of :120.
  * `cross_check` (default False = reference C++ behaviour) adds the mutual check the north star asks for.
  * prepare(frames) uploads every frame once and matches all pairs in one device pass; afterwards the
    per-pair calls are lookups.  Without prepare() each call uploads its two frames (the unmodified
    caller of cpp_code/test/sfm.cpp:153,156 works either way).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .capi import DMATCH_DTYPE, KIND_B256, KIND_F32X64, Context


@dataclass
class Frame:
    """The slice of frame_t the matching path reads: frame_id and descriptors (utility.h:23,31)."""
    frame_id: int
    descriptors: Optional[np.ndarray] = None
    keypoints: Optional[Sequence] = None
    rgb_image: Optional[np.ndarray] = None        # utility.h:25 (BGR or gray, uint8)
    image_file_path: str = ""
    unique_pixel_ids: List[int] = field(default_factory=list)


class FeatureMatching:
    def __init__(self, ctx: Optional[Context] = None, device: int = 0, cross_check: bool = False, verbose: bool = False):
        self.ctx = ctx if ctx is not None else Context(device)
        self.cross_check = bool(cross_check)
        self.verbose = bool(verbose)
        self._prepared = None  # (kind, ratio, cross_check, results, {frame_id: bank index})

    # ---- extraction (feature_matching.cpp:14-41) ----
    def detectFeaturesORB(self, cur_frame: Frame, max_num: int = 5000, show: bool = False) -> bool:
        """Key points (KEYPOINT_DTYPE records, cv2's order) and 32-byte descriptors of cur_frame.rgb_image, computed on the device."""
        cur_frame.keypoints, cur_frame.descriptors = self.ctx.orb_extract(cur_frame.rgb_image, max_num)
        if self.verbose:  # feature_matching.cpp:30 (the reference prints descriptors.size(), a cv::Size)
            print(f"Found [32 x {len(cur_frame.keypoints)}] features")
        return True

    def prepare_from_images(self, frames: Sequence[Frame], ratio_thre: float, max_num: int = 5000, cross_check: Optional[bool] = None):
        """detectFeaturesORB for every frame with the descriptors going straight into the bank (they never visit the host), then the
        all-pairs pass of prepare().  frame.keypoints is filled; frame.descriptors is left as it was."""
        cc = self.cross_check if cross_check is None else bool(cross_check)
        from .capi import Bank
        bank = Bank(self.ctx, KIND_B256, len(frames))
        for k, f in enumerate(frames):
            f.keypoints = bank.set_frame_from_image(k, f.rgb_image, max_num)
        bank.commit()
        res = bank.match_all_pairs(ratio_thre, cc)
        self._prepared = (bank.kind, float(ratio_thre), cc, res, {int(f.frame_id): k for k, f in enumerate(frames)}, bank)
        return res

    # ---- all-pairs pre-pass (the hook a maintainer inserts at cpp_code/test/sfm.cpp:131) ----
    def prepare(self, frames: Sequence[Frame], ratio_thre: float, cross_check: Optional[bool] = None):
        cc = self.cross_check if cross_check is None else bool(cross_check)
        bank = self.ctx.bank_from_frames([f.descriptors for f in frames])
        res = bank.match_all_pairs(ratio_thre, cc)
        self._prepared = (bank.kind, float(ratio_thre), cc, res, {int(f.frame_id): k for k, f in enumerate(frames)}, bank)
        return res

    def _lookup(self, kind, f1: Frame, f2: Frame, ratio: float):
        if self._prepared is None:
            return None
        pkind, pratio, pcc, res, index, _bank = self._prepared
        if pkind != kind or pratio != float(ratio) or pcc != self.cross_check:
            return None
        a, b = index.get(int(f1.frame_id)), index.get(int(f2.frame_id))
        if a is None or b is None or a <= b:
            return None
        return res.pair(a, b)

    def _match(self, kind, tag, f1: Frame, f2: Frame, matches: list, ratio_thre: float) -> bool:
        m = self._lookup(kind, f1, f2, ratio_thre)
        if m is None:
            m = self.ctx.match_descriptors(f1.descriptors, f2.descriptors, ratio_thre, self.cross_check)
        if self.verbose:  # the reference's stdout lines, feature_matching.cpp:81,96-97
            print("Initial matching done.")
            print(f"# Correspondence: Initial [ {len(f1.descriptors)} ]  Filtered by Lowe ratio test [ {len(m)} ]")
        matches.extend(m)  # push_back semantics: append, never clear (feature_matching.cpp:90)
        return True

    def matchFeaturesORB(self, cur_frame_1: Frame, cur_frame_2: Frame, matches: list, ratio_thre: float = 0.8,
                         show: bool = False) -> bool:
        d = np.asarray(cur_frame_1.descriptors)
        if d.dtype != np.uint8:
            raise TypeError("matchFeaturesORB needs uint8 (CV_8UC1) descriptors")
        return self._match(KIND_B256, "ORB", cur_frame_1, cur_frame_2, matches, ratio_thre)

    def matchFeaturesSURF(self, cur_frame_1: Frame, cur_frame_2: Frame, matches: list, ratio_thre: float = 0.5,
                          show: bool = False) -> bool:
        d = np.asarray(cur_frame_1.descriptors)
        if d.dtype != np.float32:
            raise TypeError("matchFeaturesSURF needs float32 (CV_32FC1) descriptors")
        return self._match(KIND_F32X64, "SURF", cur_frame_1, cur_frame_2, matches, ratio_thre)


def pairwise_match_descriptors(descs1, descs2, match_strategy: str, nn_ratio: float = 0.7, ctx: Optional[Context] = None):
    """The matcher half of python_code/feature_match.py:pairwise_match on caller-supplied descriptors.

    'mutual_nn'  -> :24-28  BFMatcher(NORM_L2, crossCheck=True).match, sorted by distance ascending
    'ratio_test' -> :31-39  knnMatch(k=2) + `m.distance < nn_ratio * n.distance`
    Returns a DMATCH_DTYPE array.
    """
    ctx = ctx if ctx is not None else Context(0)
    if match_strategy == "mutual_nn":
        m = ctx.match_descriptors(descs1, descs2, math.inf, True)  # ratio = +inf disables the ratio test
        order = np.argsort(m["distance"], kind="stable")
        return m[order]
    if match_strategy == "ratio_test":
        return ctx.match_descriptors(descs1, descs2, nn_ratio, False)
    raise ValueError(f"unknown match_strategy {match_strategy!r}")


__all__ = ["Frame", "FeatureMatching", "pairwise_match_descriptors", "DMATCH_DTYPE"]
