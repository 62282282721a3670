# round 2, session 4: Hamming tensor path -- full GPU suite (both paths parametrised), threshold sweep on a small problem, landmark-sized call
O=gpurun_out/r5r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
python - > $O/sizes.txt 2>&1 <<'PY'
import os, time, numpy as np, torch
from zenslam_b200.runtime import Context, match_hamming_knn2, match_hamming_cross
ctx = Context(0)
rng = np.random.default_rng(1)
def run(pairs, nq, nt, reps=30):
    q = torch.from_numpy(rng.integers(0, 256, (pairs, nq, 32), dtype=np.uint8)).cuda(); t = torch.from_numpy(rng.integers(0, 256, (pairs, nt, 32), dtype=np.uint8)).cuda()
    cq = torch.full((pairs,), nq, dtype=torch.int32, device="cuda"); ct = torch.full((pairs,), nt, dtype=torch.int32, device="cuda")
    out = {}
    for name, env in (("cuda_core", {"ZS_HAMMING_NO_TENSOR": "1"}), ("tensor", {"ZS_HAMMING_TENSOR_MIN": "1"})):
        for k in ("ZS_HAMMING_NO_TENSOR", "ZS_HAMMING_TENSOR_MIN"): os.environ.pop(k, None)
        os.environ.update(env); ctx.reload_switches()
        for _ in range(3): r = match_hamming_knn2(ctx, q, cq, t, ct, 0.8)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for _ in range(reps): r = match_hamming_knn2(ctx, q, cq, t, ct, 0.8)
        e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / reps * 1000.0; out[name + "_sum"] = int(r[0].sum())
    print(pairs, nq, nt, "work 2^%.1f" % np.log2(pairs * nq * nt), "cuda-core %.1f us" % out["cuda_core"], "tensor %.1f us" % out["tensor"], "same" if out["cuda_core_sum"] == out["tensor_sum"] else "DIFF")
for p, a, b in [(1, 500, 500), (1, 1100, 1100), (1, 2500, 2500), (2, 2500, 2500), (1, 4000, 4000), (8, 1100, 1100), (1, 1100, 30000), (1, 300, 100000), (32, 2500, 2500)]:
    run(p, a, b)
PY
cat $O/sizes.txt
