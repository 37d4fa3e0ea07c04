"""The C++ shim that keeps p3dv::FeatureMatching::matchFeaturesORB/SURF's signatures (easysfm_b200/shim) compiles against a
stand-in for the EasySFM/OpenCV headers and links against the C-ABI library (CPU); on the GPU it is driven like
cpp_code/test/sfm.cpp:140-161 and must reproduce the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "easysfm_b200", "shim", "feature_matching_gpu.cpp")
LIBDIR = os.path.join(ROOT, "easysfm_b200", "lib")


def build_shim(tmp):
    exe = os.path.join(tmp, "shim_main")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           SHIM, os.path.join(ROOT, "tests", "shim", "shim_main.cpp"), "-L", LIBDIR, "-lesfm_match", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libesfm_match.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    exe = build_shim(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("feature,prepare", [("O", 0), ("O", 1), ("S", 0), ("S", 1)])
def test_shim_matches_oracle(tmp_path, feature, prepare):
    import oracle
    from easysfm_b200 import synth
    from util import justify_l2
    exe = build_shim(str(tmp_path))
    rows = [300, 0, 257, 190]
    frames = (synth.orb_like if feature == "O" else synth.surf_like)(len(rows), rows, seed=12)
    blob = os.path.join(str(tmp_path), "desc.bin")
    with open(blob, "wb") as f:
        for fr in frames:
            f.write(np.ascontiguousarray(fr).tobytes())
    out = os.path.join(str(tmp_path), "out.bin")
    subprocess.check_call([exe, feature, str(len(rows))] + [str(r) for r in rows] + [blob, str(prepare), out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    pos = 0
    ratio = 0.8 if feature == "O" else 0.5   # the header defaults the reference's call sites rely on (feature_matching.h:17-21)
    for i in range(len(rows)):
        for j in range(i):
            cnt = int(np.frombuffer(raw, np.int32, 1, pos)[0]); pos += 4
            m = np.frombuffer(raw, oracle.DMATCH_DTYPE, cnt, pos); pos += 16 * cnt
            ref = oracle.match(frames[i], frames[j], ratio, False)
            if feature == "O":
                assert len(m) == len(ref) and (m["trainIdx"] == ref["trainIdx"]).all() and (m["distance"] == ref["distance"]).all()
                assert (m["imgIdx"] == 0).all()
            elif len(frames[i]) and len(frames[j]):
                justify_l2(frames[i], frames[j], ratio, False, m, ref)
            else:
                assert cnt == 0
    assert pos == len(raw)
