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

SOURCES = ["context.cpp", "pyr_lk.cpp", "pyr_lk_factory.cpp", "keypoint_detector_cuda.cpp", "bf_matcher.cpp", "stereo_tracker.cpp", "processing.cpp"]


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


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not present")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_adapter_links_and_falls_back_without_a_device(tmp_path):
    """The adapter + harness COMPILE AND LINK against libzenslam_cuda.so (tests/adapter/build_harness.py).  On a machine
    without an sm_100 device create_cuda_pyr_lk() returns an empty pointer -- the contract of the Metal factory it sits
    beside (zenslam_metal/source/pyr_lk_factory.cpp:41-49) -- and the harness reports exactly that (exit code 3); with a
    B200 it runs to the end (tests/test_gpu_adapter.py)."""
    import sys

    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "adapter"))
    import build_harness
    exe = build_harness.build()
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(b"dims i4 1 7\n" + np.array([32, 32, 1, 16, 10, 31, 3], np.int32).tobytes() + b"\n")
        f.write(b"frames u1 4 1 2 32 32\n" + bytes(2 * 32 * 32) + b"\n")
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert r.returncode != 3, r.stderr
    else:
        assert r.returncode == 3 and "empty pointer" in r.stderr, (r.returncode, r.stderr)
