# round 2, GPU call v: 63x63 KLT with two tiles per warp (ZS_KLT63_TWO_TILES = CTAs per SM) -- parity, then A/B
mkdir -p gpurun_out/r2v && O=gpurun_out/r2v
for mb in 5; do
ZS_KLT63_TWO_TILES=$mb timeout 900 python -m pytest tests -m gpu -q -x -k "lk or klt or tumvi or fullsize or frontend_batches or properties" > $O/pytest_$mb.log 2>&1; echo "pytest($mb) exit $?" | tee -a $O/pytest_$mb.log; tail -3 $O/pytest_$mb.log
done
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --config TUMVI --steps 5"
timeout 300 $B > $O/tumvi_v4.json 2> $O/tumvi_v4.err
for mb in 4 5 6; do ZS_KLT63_TWO_TILES=$mb timeout 300 $B > $O/tumvi_v5_$mb.json 2> $O/tumvi_v5_$mb.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), d["stage_ms_per_step"]["klt"])
    except Exception as e: print(f, "ERR", e)
PY
