"""Build tests/adapter/_build/adapter_harness: the zenslam_cuda/ C++ adapter + tests/adapter/adapter_harness.cpp, compiled
against the reference's own headers (read where they lie under /root/reference, never copied) and the functional OpenCV
stand-in under tests/stubs/, linked to the in-tree libzenslam_cuda.so.  /root/reference only exists in the build
container, so the executable is built here (by __graft_entry__.build()) and travels to the GPU box with the snapshot,
like the .so files; tests/test_gpu_adapter.py runs it there.

    python tests/adapter/build_harness.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INC = "/root/reference/zenslam_core/include"
OUT = os.path.join(HERE, "_build")
EXE = os.path.join(OUT, "adapter_harness")
ADAPTER = os.path.join(ROOT, "zenslam_cuda")
SOURCES = ["context.cpp", "pyr_lk.cpp", "pyr_lk_factory.cpp", "keypoint_detector_cuda.cpp", "bf_matcher.cpp", "stereo_tracker.cpp", "processing.cpp"]


def available() -> bool:
    return os.path.isdir(REF_INC)


def build(force: bool = False) -> str:
    """-> path of the executable; raises when the reference headers are absent or g++ fails"""
    if not available():
        raise RuntimeError("reference headers not present (%s): the harness is built in the container that has them" % REF_INC)
    from zenslam_b200.build import LIB, build as build_cuda
    build_cuda()
    srcs = [os.path.join(HERE, "adapter_harness.cpp")] + [os.path.join(ADAPTER, "source", s) for s in SOURCES]
    deps = srcs + [LIB, os.path.abspath(__file__)]
    for d, _, fs in list(os.walk(os.path.join(ROOT, "tests", "stubs"))) + list(os.walk(os.path.join(ADAPTER, "include"))) + \
            list(os.walk(os.path.join(ROOT, "include"))):
        deps += [os.path.join(d, f) for f in fs]
    if not force and os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(p) for p in deps):
        return EXE
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-std=c++23", "-O1", "-g", "-Wall", "-Wextra",
           "-include", os.path.join(ROOT, "tests", "stubs", "ranges_to_shim.h"),
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ADAPTER, "include"), "-I", os.path.join(ADAPTER, "source"),
           "-I", os.path.join(ROOT, "include"), "-I", REF_INC, *srcs,
           "-L", os.path.dirname(LIB), "-lzenslam_cuda", "-Wl,-rpath,$ORIGIN/../../../zenslam_b200/_build", "-lpthread",
           "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("adapter harness build failed:\n" + r.stderr[-6000:])
    return EXE


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build(force="--force" in sys.argv))
