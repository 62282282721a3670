# round 2, GPU call q: adapter with landmarks, keyline matching, subpix terms loop -- full GPU suite + TUMVI line
mkdir -p gpurun_out/r2q && O=gpurun_out/r2q
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -12 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline"
timeout 300 $B --config TUMVI --steps 5 > $O/tumvi.json 2> $O/tumvi.err
timeout 300 $B --config TUMVI752 --steps 10 > $O/tumvi752.json 2> $O/tumvi752.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["stage_ms_per_step"]["fast_grid"])
    except Exception as e: print(f, "ERR", e)
PY
