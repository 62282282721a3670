# round 2, session 4: packed 63x63 loop with the sample left at acc * 128 (IMAD) and PRMT on bytes 2, 3 instead of >> 9 (SHF)
O=gpurun_out/r5k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_random_sweep.py tests/test_gpu_fullsize.py -m gpu -x -q -k "63 or fullsize" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3 --config TUMVI"
timeout 300 $B > $O/tumvi.json 2> $O/tumvi.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5k/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
