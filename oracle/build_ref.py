"""oracle/_ref/libzs_ref_glue.so -- the REFERENCE's own detector glue, run here.

The three detector sources (zenslam_core/source/detection/keypoint_detector_{grid,parallel,simple}.cpp) are compiled UNMODIFIED
from where they lie under /root/reference, with the reference's own class headers, against
  * the functional OpenCV stand-in of tests/stubs/ (the image has no C++ OpenCV),
  * oracle/ref_glue/opencv_over_oracle.cpp: cv::FastFeatureDetector / cv::ORB / cv::cornerSubPix implemented over the C oracle
    (whose arithmetic is pinned to real cv2 outputs), and
  * four shadow headers (oracle/ref_glue/shadow/): zenslam/all_options.h and zenslam/utils/utils_opencv.h (the real ones pull
    yaml-cpp, the IMU integrator, viz and frame types in; the detectors use detection_options and one cv::Size operator from
    them), gsl/narrow, opencv2/xfeatures2d.hpp.
Nothing is copied from the reference; the output lands in oracle/_ref/ (git-ignored, travels to the GPU box like the other
built objects).  tests/test_oracle_vs_reference_glue.py compares the oracle's restatement of the glue with it.

    python oracle/build_ref.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/zenslam_core"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libzs_ref_glue.so")
REF_SOURCES = ["source/detection/keypoint_detector_grid.cpp", "source/detection/keypoint_detector_parallel.cpp",
               "source/detection/keypoint_detector_simple.cpp"]


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, s)) for s in REF_SOURCES)


def build(force: bool = False) -> str:
    if not available():
        raise RuntimeError("reference sources not present under %s" % REF)
    sys.path.insert(0, ROOT)
    import oracle
    oracle_so = oracle.build()
    glue = os.path.join(HERE, "ref_glue")
    srcs = [os.path.join(REF, s) for s in REF_SOURCES] + [os.path.join(glue, "opencv_over_oracle.cpp"), os.path.join(glue, "ref_glue.cpp")]
    deps = list(srcs) + [oracle_so, os.path.abspath(__file__)]
    for d in (os.path.join(glue, "shadow"), os.path.join(ROOT, "tests", "stubs")):
        for dd, _, fs in os.walk(d):
            deps += [os.path.join(dd, f) for f in fs]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-std=c++23", "-O1", "-g", "-shared", "-fPIC", "-fvisibility=hidden",
           "-include", os.path.join(ROOT, "tests", "stubs", "ranges_to_shim.h"),
           "-I", os.path.join(glue, "shadow"),                   # shadows first: all_options.h, utils_opencv.h, gsl, xfeatures2d
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(REF, "include"),
           *srcs, "-L", os.path.dirname(oracle_so), "-lzs_oracle", "-Wl,-rpath,$ORIGIN/../_build", "-lpthread", "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference glue build failed:\n" + r.stderr[-8000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
