# round 2, session 4: grid FAST with one cell per block folds the cell maximum per warp (REDUX) before the shared-memory atomic
O=gpurun_out/r5h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "fast or grid or detect or tumvi or TUMVI or parallel or frontend" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B --config TUMVI > $O/tumvi.json 2> $O/tumvi.err
timeout 300 $B --config TUMVI752 > $O/tumvi752.json 2> $O/tumvi752.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5h/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()}, d.get("detect_ms"))
    except Exception as e: print(f, "ERR", e)
PY
