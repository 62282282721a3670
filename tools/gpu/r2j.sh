# round 2, GPU call j: persistent KLT launch -- parity, then A/B at C2 / TUMVI / small batches / tracker
mkdir -p gpurun_out/r2j && O=gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --steps 10"
timeout 300 $B > $O/c2_persist.json 2> $O/c2_persist.err
ZS_KLT_NO_PERSIST=1 timeout 300 $B > $O/c2_percta.json 2> $O/c2_percta.err
timeout 300 $B --config TUMVI --steps 5 > $O/tumvi_persist.json 2> $O/tumvi_persist.err
ZS_KLT_NO_PERSIST=1 timeout 300 $B --config TUMVI --steps 5 > $O/tumvi_percta.json 2> $O/tumvi_percta.err
timeout 300 $B --batch 16 --steps 30 > $O/b16_persist.json 2> $O/b16_persist.err
ZS_KLT_NO_PERSIST=1 timeout 300 $B --batch 16 --steps 30 > $O/b16_percta.json 2> $O/b16_percta.err
timeout 300 $B --batch 1 --steps 50 > $O/b1_persist.json 2> $O/b1_persist.err
ZS_KLT_NO_PERSIST=1 timeout 300 $B --batch 1 --steps 50 > $O/b1_percta.json 2> $O/b1_percta.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["stage_ms_per_step"]["klt"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python tools/bench_tracker.py > $O/tracker_persist.json 2>&1
ZS_KLT_NO_PERSIST=1 timeout 300 python tools/bench_tracker.py > $O/tracker_percta.json 2>&1
tail -c 700 $O/tracker_persist.json; echo; tail -c 700 $O/tracker_percta.json
