"""The options.yaml keys the hot path reads, with the reference's names, nesting and defaults.

Mirrors zenslam_core/include/zenslam/detection/detection_options.h:11-23,
zenslam_core/include/zenslam/tracking_options.h:7-18 and slam_options
(zenslam_core/include/zenslam/all_options.h:111-137); parsing follows options_parser.cpp:66-127,239-283:
a missing or invalid key keeps its default, sizes are 2-element sequences, the detector key is
``feature`` (the writer's ``feature_detector`` is ignored, SURVEY Appendix B.2 -- reproduced, not fixed).
Only the keys of the hot path are modelled; everything else in the file is ignored.
"""
from __future__ import annotations

from dataclasses import dataclass, field

FEATURE_TYPES = ("FAST", "ORB", "SIFT")                  # detection/feature_type.h:5-10
DESCRIPTOR_TYPES = ("ORB", "SIFT", "FREAK")              # detection/descriptor_type.h:5-10
DETECTION_ALGORITHMS = ("SIMPLE", "GRID", "PARALLEL_GRID")   # detection/detection_algorithm.h:5-10
MATCHER_TYPES = ("BRUTE", "KNN", "FLANN")


@dataclass
class detection_options:
    clahe_enabled: bool = False
    cell_size: tuple = (16, 16)
    fast_threshold: int = 10
    feature_detector: str = "FAST"
    descriptor: str = "ORB"
    algorithm: str = "GRID"


@dataclass
class tracking_options:
    klt_window_size: tuple = (31, 31)
    klt_max_level: int = 3
    klt_threshold: float = 1.0
    klt_min_tracked_ratio: float = 0.6
    landmark_match_distance: float = 32.0
    landmark_match_radius: float = 50.0
    filter_epipolar: bool = True
    epipolar_threshold: float = 1.0


@dataclass
class slam_options:
    matcher: str = "BRUTE"
    matcher_ratio: float = 0.8
    epipolar_threshold: float = 1.0
    detection: detection_options = field(default_factory=detection_options)
    tracking: tracking_options = field(default_factory=tracking_options)


def _get(node, key, default, cast):
    if not isinstance(node, dict) or key not in node:
        return default
    try:
        return cast(node[key])
    except (TypeError, ValueError):
        return default


def _size(node, key, default):
    if not isinstance(node, dict) or key not in node:
        return default
    v = node[key]
    try:
        if isinstance(v, (list, tuple)) and len(v) >= 2:
            return (int(v[0]), int(v[1]))
    except (TypeError, ValueError):
        pass
    return default


def _enum(node, key, default, allowed):
    if isinstance(node, dict) and key in node and str(node[key]) in allowed:
        return str(node[key])
    return default


def parse_slam(node) -> slam_options:
    """options_parser::parse_slam (options_parser.cpp:239-283) for the hot-path keys."""
    o = slam_options()
    if not isinstance(node, dict):
        return o
    o.matcher = _enum(node, "matcher", o.matcher, MATCHER_TYPES)
    o.matcher_ratio = _get(node, "matcher_ratio", o.matcher_ratio, float)
    o.epipolar_threshold = _get(node, "epipolar_threshold", o.epipolar_threshold, float)
    d, dn = o.detection, node.get("detection")
    d.clahe_enabled = _get(dn, "clahe_enabled", d.clahe_enabled, bool)
    d.cell_size = _size(dn, "cell_size", d.cell_size)
    d.fast_threshold = _get(dn, "fast_threshold", d.fast_threshold, int)
    d.feature_detector = _enum(dn, "feature", d.feature_detector, FEATURE_TYPES)
    d.descriptor = _enum(dn, "descriptor", d.descriptor, DESCRIPTOR_TYPES)
    d.algorithm = _enum(dn, "algorithm", d.algorithm, DETECTION_ALGORITHMS)
    t, tn = o.tracking, node.get("tracking")
    t.klt_window_size = _size(tn, "klt_window_size", t.klt_window_size)
    t.klt_max_level = _get(tn, "klt_max_level", t.klt_max_level, int)
    t.klt_threshold = _get(tn, "klt_threshold", t.klt_threshold, float)
    t.klt_min_tracked_ratio = _get(tn, "klt_min_tracked_ratio", t.klt_min_tracked_ratio, float)
    t.landmark_match_distance = _get(tn, "landmark_match_distance", t.landmark_match_distance, float)
    t.landmark_match_radius = _get(tn, "landmark_match_radius", t.landmark_match_radius, float)
    t.filter_epipolar = _get(tn, "filter_epipolar", t.filter_epipolar, bool)
    t.epipolar_threshold = _get(tn, "epipolar_threshold", t.epipolar_threshold, float)
    return o


def load(path: str) -> slam_options:
    """options_parser::load for the `slam:` subtree of an options.yaml."""
    import yaml
    with open(path) as f:
        root = yaml.safe_load(f)
    return parse_slam(root.get("slam") if isinstance(root, dict) else None)
