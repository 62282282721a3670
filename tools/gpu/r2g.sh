# round 2, GPU call g: cornerSubPix v2 tuned -- parity tests, TUMVI A/B (ZS_SUBPIX_V1 set = old kernel), ncu
mkdir -p gpurun_out/r2g && O=gpurun_out/r2g
timeout 1200 python -m pytest tests -m gpu -q -x -k "subpix or parallel or tumvi or fullsize or adapter" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 300 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi_v2.json 2> $O/bench_tumvi_v2.err
ZS_SUBPIX_V1=1 timeout 300 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi_v1.json 2> $O/bench_tumvi_v1.err
timeout 300 python bench.py --config TUMVI752 --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi752_v2.json 2> $O/bench_tumvi752_v2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["stage_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
TV="python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corner_subpix_v2 -s 8 -c 1 -o $O/subpix_v2 $TV > $O/ncu_subpix.log 2>&1
