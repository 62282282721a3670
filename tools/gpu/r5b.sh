# round 2, session 4: C3 with the double-buffered end-to-end step
O=gpurun_out/r5b; mkdir -p $O
timeout 600 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > $O/c3.json 2> $O/c3.err; echo "c3 exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5b/c3.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "u8", d["u8_rows"])
PY
tail -3 $O/c3.err
