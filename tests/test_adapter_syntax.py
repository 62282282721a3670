"""The C++ adapter in zenslam_cuda/ cannot be linked here (no C++ OpenCV in the image), but it can be type-checked:
g++ -fsyntax-only against the REFERENCE's own headers (pyr_lk.h, keypoint_detector.h, detection_options.h) plus a
test-only stand-in for the handful of OpenCV/spdlog declarations they use (tests/stubs/).  This proves the adapter
classes really override the reference's virtual interfaces with the exact signatures.  Skipped where
/root/reference is absent (the GPU box)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/zenslam_core/include"

SOURCES = ["context.cpp", "pyr_lk.cpp", "pyr_lk_factory.cpp", "keypoint_detector_cuda.cpp", "bf_matcher.cpp", "stereo_tracker.cpp"]


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not present")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.parametrize("src", SOURCES)
def test_adapter_type_checks_against_reference_headers(src):
    cmd = ["g++", "-std=c++23", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=overloaded-virtual",
           "-include", os.path.join(ROOT, "tests", "stubs", "ranges_to_shim.h"),
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "zenslam_cuda", "include"),
           "-I", os.path.join(ROOT, "include"), "-I", REF_INC, os.path.join(ROOT, "zenslam_cuda", "source", src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
