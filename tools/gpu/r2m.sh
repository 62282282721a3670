# round 2, GPU call m: subpix v2b with odd term stride; property tests; TUMVI both sizes
mkdir -p gpurun_out/r2m && O=gpurun_out/r2m
timeout 900 python -m pytest tests -m gpu -q -k "properties or subpix or parallel or tumvi or fullsize" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline"
for i in 1 2; do
timeout 300 $B --config TUMVI --steps 5 > $O/tumvi_$i.json 2> $O/tumvi.err
timeout 300 $B --config TUMVI752 --steps 10 > $O/tumvi752_$i.json 2> $O/tumvi752.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["stage_ms_per_step"]["fast_grid"])
    except Exception as e: print(f, "ERR", e)
PY
TV="python bench.py --config TUMVI752 --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corner_subpix_v2 -s 8 -c 1 -o $O/subpix_752 $TV > $O/ncu_subpix.log 2>&1
