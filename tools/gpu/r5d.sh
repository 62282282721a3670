# round 2, session 4: packed-template form of the 31x31 KLT kernel (ZS_KLT31_PACKED = CTAs per SM) against the shipped one
O=gpurun_out/r5d; mkdir -p $O
ZS_KLT31_PACKED=28 timeout 600 python -m pytest tests/test_gpu_random_sweep.py tests/test_gpu_parity.py tests/test_gpu_frontend.py -m gpu -x -q -k "klt or frontend" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B > $O/c2_default.json 2> $O/c2_default.err
for m in 24 28 32; do ZS_KLT31_PACKED=$m timeout 300 $B > $O/c2_packed$m.json 2> $O/c2_packed$m.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5d/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
