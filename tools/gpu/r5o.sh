# round 2, session 4: same-pixel fast path in the 31x31 Gauss-Newton loop (bounds / re-fetch tests and patch offsets only when the integer origin moves)
O=gpurun_out/r5o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_random_sweep.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "klt or fullsize" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 10 --warmup 3"
timeout 300 $B > $O/c2.json 2> $O/c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5o/c2.json").read().strip().splitlines()[-1]); print(round(d["value"]), round(d["e2e"]["value"]), {k:round(v,3) for k,v in d.get("stage_ms_per_step").items()})
PY
