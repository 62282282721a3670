"""CPU checks of the drop-in boundary: the library loads without a GPU, exports every symbol the header
declares, and refuses to compute without a device (no fallback)."""
import ctypes as C
import os
import re

import pytest

from zenslam_b200 import _lib
from zenslam_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build()
    return _lib.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "zenslam_cuda.h")).read()
    return sorted(set(re.findall(r"ZS_API\s+[\w\s\*]+?\b(zs_\w+)\s*\(", src)))


def test_header_symbols_all_exported(L):
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), "libzenslam_cuda.so does not export " + s


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.zs_is_available() == 0
    h = C.c_void_p()
    assert L.zs_context_create(0, None, C.byref(h)) == -1       # ZS_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.zs_last_error_string()
    from zenslam_b200.runtime import Context
    with pytest.raises(_lib.ZenslamCudaError):
        Context()
    from zenslam_b200.tracking import create_cuda_pyr_lk
    assert create_cuda_pyr_lk() is None                          # factory contract of pyr_lk_factory.cpp:41-49


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "zenslam_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "zs_oracle" not in txt, fn


def test_options_parser_keys(tmp_path):
    from zenslam_b200 import options
    y = tmp_path / "o.yaml"
    y.write_text("slam:\n  matcher: KNN\n  matcher_ratio: 0.7\n  detection:\n    cell_size: [64, 64]\n"
                 "    fast_threshold: 1\n    feature_detector: ORB\n    algorithm: PARALLEL_GRID\n"
                 "  tracking:\n    klt_window_size: [63, 63]\n    klt_max_level: 4\n    klt_threshold: 2\n")
    o = options.load(str(y))
    assert o.matcher == "KNN" and o.matcher_ratio == 0.7
    assert o.detection.cell_size == (64, 64) and o.detection.fast_threshold == 1
    assert o.detection.feature_detector == "FAST"      # parser reads `feature`, not `feature_detector` (Appendix B.2)
    assert o.detection.algorithm == "PARALLEL_GRID"
    assert o.tracking.klt_window_size == (63, 63) and o.tracking.klt_max_level == 4 and o.tracking.klt_threshold == 2.0
    d = options.slam_options()
    assert d.matcher == "BRUTE" and d.detection.cell_size == (16, 16) and d.tracking.klt_window_size == (31, 31)
