# round 2, session 4: two-copy exact-origin J staging of the tiled KLT kernel (JC = 2) against the one-copy form
O=gpurun_out/r5a; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B > $O/c2_two_copy.json 2> $O/c2_two_copy.err; echo "c2 exit $?"
ZS_KLT_ONE_COPY=1 timeout 300 $B > $O/c2_one_copy.json 2> $O/c2_one_copy.err
ZS_KLT63_FOUR_WARPS=1 timeout 300 $B --config TUMVI > $O/tumvi_4w_two_copy.json 2> $O/tumvi_4w_two_copy.err
ZS_KLT63_FOUR_WARPS=1 ZS_KLT_ONE_COPY=1 timeout 300 $B --config TUMVI > $O/tumvi_4w_one_copy.json 2> $O/tumvi_4w_one_copy.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5a/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], d["value"], d["e2e"]["value"], d.get("stage_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
