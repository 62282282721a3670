"""zenslam_b200 -- B200-native (sm_100a) backend for ZenSLAM's stereo front-end hot path.

Everything numerical lives in libzenslam_cuda.so (zenslam_b200/csrc, C ABI in include/zenslam_cuda.h);
this package is the host-side mirror of the reference's detector / matcher / pyr_lk interfaces.
There is no CPU fallback: importing works anywhere, but creating a Context without a B200 raises.
"""
from ._lib import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW, ZenslamCudaError, lib  # noqa: F401
from .options import detection_options, slam_options, tracking_options  # noqa: F401
from .types import DMatch, keyline, keypoint  # noqa: F401

__all__ = ["lib", "ZenslamCudaError", "slam_options", "detection_options", "tracking_options", "keypoint", "DMatch",
           "LK_GET_MIN_EIGENVALS", "LK_USE_INITIAL_FLOW"]
