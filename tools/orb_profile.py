"""Two ORB extractions of one 1080p frame, for an ncu launch list (tools: ncu --metrics gpu__time_duration.sum ... python tools/orb_profile.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from easysfm_b200 import capi  # noqa: E402
from orb_util import image  # noqa: E402

ctx = capi.Context(0)
im = image(900, 1080, 1920, 400, bgr=True)
for _ in range(2):
    kp, d = ctx.orb_extract(im, 5000)
print(len(kp), ctx.orb_last_timing())
