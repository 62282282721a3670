# round 2, GPU call f: cornerSubPix v2 (six points per warp) -- parity tests, then TUMVI A/B
mkdir -p gpurun_out/r2f && O=gpurun_out/r2f
timeout 1200 python -m pytest tests -m gpu -q -x -k "subpix or parallel or tumvi or fullsize or landmarks or adapter or errors" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
for v in 0 1; do
ZS_SUBPIX_V1=$v timeout 300 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi_v1_$v.json 2> $O/bench_tumvi_v1_$v.err
ZS_SUBPIX_V1=$v timeout 300 python bench.py --config TUMVI752 --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi752_v1_$v.json 2> $O/bench_tumvi752_v1_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["stage_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
TV="python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corner_subpix_v2 -s 8 -c 1 -o $O/subpix_v2 $TV > $O/ncu_subpix.log 2>&1
