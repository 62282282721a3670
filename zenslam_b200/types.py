"""Host data model crossing the boundary (zenslam_core/include/zenslam/types/keypoint.h:8-15)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class keypoint:
    """zenslam::keypoint : cv::KeyPoint {size_t index; cv::Mat descriptor; static index_next}."""
    pt: tuple = (0.0, 0.0)
    size: float = 7.0
    angle: float = -1.0
    response: float = 0.0
    octave: int = 0
    class_id: int = -1
    index: int = 0
    descriptor: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.uint8))

    index_next = 0      # class-level counter (types/keypoint.cpp:3); not thread-safe in the reference either


@dataclass
class keyline:
    """zenslam::keyline : cv::line_descriptor::KeyLine {size_t index; ...} (types/keyline.h:9) -- the fields
    utils::track_keylines reads and rewrites (tracking_utils.cpp:30-36,121-137)."""
    startPointX: float = 0.0
    startPointY: float = 0.0
    endPointX: float = 0.0
    endPointY: float = 0.0
    pt: tuple = (0.0, 0.0)
    lineLength: float = 0.0
    angle: float = 0.0
    octave: int = 0
    class_id: int = -1
    index: int = 0
    descriptor: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.uint8))


@dataclass
class DMatch:
    """cv::DMatch as matcher::match_keypoints returns it: indices are KEYPOINT indices (matcher.cpp:110,213)."""
    queryIdx: int
    trainIdx: int
    distance: float
