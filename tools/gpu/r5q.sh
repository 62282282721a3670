# round 2, session 4: Hamming kNN on the tcgen05 kind::i8 kernel (bits expanded to bytes, two K halves)
O=gpurun_out/r5q; mkdir -p $O
ZS_HAMMING_TENSOR_MIN=1 timeout 600 python -m pytest tests -m gpu -x -q -k "hamming or match or landmark or knn or cross" > $O/pytest_forced.log 2>&1; echo "pytest exit $?" >> $O/pytest_forced.log; tail -5 $O/pytest_forced.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B > $O/c2_tensor.json 2> $O/c2_tensor.err
ZS_HAMMING_NO_TENSOR=1 timeout 300 $B > $O/c2_cuda_core.json 2> $O/c2_cuda_core.err
timeout 300 $B --config C5 --steps 5 > $O/c5_tensor.json 2> $O/c5_tensor.err
ZS_HAMMING_NO_TENSOR=1 timeout 300 $B --config C5 --steps 5 > $O/c5_cuda_core.json 2> $O/c5_cuda_core.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5q/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -2 $O/c2_tensor.err
