# round 2, GPU call w: 63x63 two-tiles-per-warp at 6 / 7 / 8 CTAs per SM vs the four-warp form (two rounds)
mkdir -p gpurun_out/r2w && O=gpurun_out/r2w
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --config TUMVI --steps 5"
for r in 1 2; do
timeout 300 $B > $O/tumvi_v4_$r.json 2> $O/err.txt
for mb in 6 7 8; do ZS_KLT63_TWO_TILES=$mb timeout 300 $B > $O/tumvi_v5_${mb}_$r.json 2> $O/err.txt; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), d["stage_ms_per_step"]["klt"])
    except Exception as e: print(f, "ERR", e)
PY
