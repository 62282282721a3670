# round 2, session 4: packed-template form of the two-tile 63x63 KLT kernel (ZS_KLT63_PACKED = CTAs per SM)
O=gpurun_out/r5c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_random_sweep.py -m gpu -x -q -k "klt_random_63" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3 --config TUMVI"
timeout 300 $B > $O/tumvi_default.json 2> $O/tumvi_default.err
for m in 6 7 8; do ZS_KLT63_PACKED=$m timeout 300 $B > $O/tumvi_packed$m.json 2> $O/tumvi_packed$m.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5c/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
