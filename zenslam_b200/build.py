"""Build libzenslam_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m zenslam_b200.build [--force]

Objects and the library land in zenslam_b200/_build/ (git-ignored; shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libzenslam_cuda.so")

SOURCES = ["zs_context.cu", "zs_landmarks.cu", "zs_pyramid.cu", "zs_fast.cu", "zs_orb.cu", "zs_orb_detect.cu", "zs_match.cu", "zs_match_l2.cu",
           "zs_klt.cu", "zs_subpix.cu", "zs_preproc.cu", "zs_triangulate.cu", "zs_host.cu", "zs_frontend.cu", "zs_tracker.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                  # bit-exact float paths: FMAs only where written as fmaf()
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, hdr_mtime: float) -> str:
    obj = os.path.join(OUT, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime):
        return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    hdr_mtime = _deps()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_mtime), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        tmp = LIB + ".tmp.%d" % os.getpid()
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
               "-Xlinker", "--no-undefined", *objs, "-o", tmp, "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
        os.replace(tmp, LIB)
    if verbose:
        print(LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
