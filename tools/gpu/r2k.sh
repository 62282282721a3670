# round 2, GPU call k: subpix v2b + property tests + persist threshold: full GPU suite, TUMVI / C2 lines
mkdir -p gpurun_out/r2k && O=gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline"
timeout 300 $B --config TUMVI --steps 5 > $O/tumvi.json 2> $O/tumvi.err
timeout 300 $B --config TUMVI752 --steps 5 > $O/tumvi752.json 2> $O/tumvi752.err
timeout 300 $B --steps 10 > $O/c2.json 2> $O/c2.err
timeout 300 $B --batch 1 --steps 50 > $O/b1.json 2> $O/b1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["stage_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
