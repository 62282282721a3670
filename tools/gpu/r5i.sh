# round 2, session 4: the >> 9 of the bilinear sample as IMAD.HI (FMA pipe) instead of SHF (ALU pipe), both tiled KLT kernels
O=gpurun_out/r5i; mkdir -p $O
ZS_KLT31_PACKED=3 ZS_KLT63_PACKED=3 timeout 900 python -m pytest tests/test_gpu_random_sweep.py tests/test_gpu_parity.py -m gpu -x -q -k "klt" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B > $O/c2_shf.json 2> $O/c2_shf.err
ZS_KLT31_PACKED=3 timeout 300 $B > $O/c2_imadhi.json 2> $O/c2_imadhi.err
timeout 300 $B --config TUMVI > $O/tumvi_shf.json 2> $O/tumvi_shf.err
ZS_KLT63_PACKED=3 timeout 300 $B --config TUMVI > $O/tumvi_imadhi.json 2> $O/tumvi_imadhi.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5i/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
    except Exception as e: print(f, "ERR", e)
PY
