# round 2, GPU call t: TMA-staged grid-FAST strips -- full GPU suite, A/B at C2 / C4 / TUMVI, ncu of the kernel
mkdir -p gpurun_out/r2t && O=gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
B="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline"
for cfg in C2 C4 TUMVI; do
timeout 300 $B --config $cfg --steps 10 > $O/${cfg}_tma.json 2> $O/${cfg}_tma.err
ZS_FAST_NO_TMA=1 timeout 300 $B --config $cfg --steps 10 > $O/${cfg}_ld.json 2> $O/${cfg}_ld.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), d["stage_ms_per_step"]["fast_grid"])
    except Exception as e: print(f, "ERR", e)
PY
C2="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none -k regex:k_fast_grid_v2 -s 4 -c 1 -o $O/fast_tma $C2 > $O/ncu_fast.log 2>&1
