# round 2, session 4: packed-template 63x63 KLT as the default -- full GPU suite, TUMVI lines, ncu of the kernel
mkdir -p gpurun_out/r5e && O=gpurun_out/r5e
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline"
timeout 300 $B --config TUMVI --steps 10 > $O/tumvi.json 2> $O/err.txt
timeout 300 $B --config TUMVI752 --steps 10 > $O/tumvi752.json 2> $O/err.txt
ZS_KLT63_UNPACKED=1 timeout 300 $B --config TUMVI --steps 10 > $O/tumvi_unpacked.json 2> $O/err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5e/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["stage_ms_per_step"]["klt"])
    except Exception as e: print(f, "ERR", e)
PY
TV="python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_klt_track_v4 -s 4 -c 1 -o $O/klt63_packed $TV > $O/ncu_klt63.log 2>&1
ls -la $O
