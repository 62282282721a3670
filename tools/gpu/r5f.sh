# round 2, session 4: rolled template pass of the packed two-tile 63x63 KLT kernel (default) against the unrolled one (ZS_KLT63_PACKED=8)
O=gpurun_out/r5f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_random_sweep.py tests/test_gpu_fullsize.py tests/test_gpu_frontend.py -m gpu -x -q -k "63 or tumvi or fullsize or TUMVI" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3 --config TUMVI"
timeout 300 $B > $O/tumvi_rolled.json 2> $O/tumvi_rolled.err
ZS_KLT63_PACKED=8 timeout 300 $B > $O/tumvi_unrolled.json 2> $O/tumvi_unrolled.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5f/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
