"""easysfm_b200 -- B200-native (sm_100a) all-pairs descriptor matching behind EasySFM's matching entry point.

Only what the hot path needs lives here:
  csrc/                 CUDA kernels + the C ABI (include/esfm_match.h) -> lib/libesfm_match.so
  capi.py               ctypes binding of that C ABI (no torch types cross it)
  feature_matching.py   host-side mirror of the reference interface
                        (p3dv::FeatureMatching::matchFeaturesORB/SURF, python pairwise_match)
  scheduler.py          the all-pairs loop of cpp_code/test/sfm.cpp:140-161 as a sharded pair schedule
  synth.py              seeded synthetic descriptor banks of the BASELINE.json shapes

There is no CPU fallback anywhere in this package: without the compiled CUDA library and a B200
every compute call raises.
"""
from .capi import (  # noqa: F401
    DMATCH_DTYPE,
    KEEP_DIGESTS,
    KEEP_MATCHES,
    KIND_B256,
    KIND_F32X64,
    Bank,
    Context,
    EsfmError,
    MultiBank,
    MultiContext,
    Results,
    Tracks,
    library_path,
    load_library,
    load_results,
    read_match_file,
    write_match_file,
)
from .feature_matching import FeatureMatching, Frame, pairwise_match_descriptors  # noqa: F401
from .scheduler import all_pairs, match_all_pairs, shard_pairs  # noqa: F401
from .motion import MotionEstimator  # noqa: F401

__version__ = "0.2.0"
