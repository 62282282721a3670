#!/usr/bin/env python
"""Micro-benchmark of the L2 (SIFT) matcher at BASELINE config 3 shapes: 2000 x 2000 x 128, kNN-2 + ratio, `pairs`
independent problems per launch.  Reports time, effective integer tensor throughput (2*Nq*Nt*128 ops per direction)
against the measured dense bf16 peak, and the CUDA-core dp4a kernel beside it.

    python tools/bench_l2.py [--pairs 64] [--n 2000] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from zenslam_b200.runtime import Context, match_l2_knn2
    ctx = Context(0)
    rng = np.random.default_rng(3)
    g = rng.gamma(0.6, 40.0, (2, a.pairs, a.n, 128))
    g = g / np.linalg.norm(g, axis=-1, keepdims=True) * 512
    d = np.clip(np.rint(g), 0, 255).astype(np.float32)
    q, t = torch.from_numpy(d[0]).cuda(), torch.from_numpy(d[1]).cuda()
    n = torch.full((a.pairs,), a.n, dtype=torch.int32, device="cuda")
    peak = 1645.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    out = {}
    for name, env in (("tcgen05_i8", None), ("cuda_core_dp4a", "1")):
        if env:
            os.environ["ZS_L2_NO_TENSOR"] = env; ctx.reload_switches()
        else:
            os.environ.pop("ZS_L2_NO_TENSOR", None); ctx.reload_switches()
        for _ in range(3):
            r = match_l2_knn2(ctx, q, n, t, n, 0.8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            r = match_l2_knn2(ctx, q, n, t, n, 0.8)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        ops = 2.0 * a.pairs * a.n * a.n * 128
        out[name] = {"ms_per_call": ms, "tera_ops_per_s": ops / ms / 1e9, "frac_of_bf16_peak": ops / ms / 1e9 / peak,
                     "idx_checksum": int(r[0].to(torch.int64).sum().item())}
    out["config"] = {"pairs": a.pairs, "n": a.n, "dim": 128, "peak_bf16_tflops": peak,
                     "note": "whole call timed: f32->u8 conversion with the row norms and the integrality flag in one pass (no host "
                             "sync), top-2 kernel, merge and ratio epilogue"}
    assert out["tcgen05_i8"]["idx_checksum"] == out["cuda_core_dp4a"]["idx_checksum"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
