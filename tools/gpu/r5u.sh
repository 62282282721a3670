# round 2, session 4: 30-second sustained run of the last tree
mkdir -p gpurun_out/r5u && O=gpurun_out/r5u
timeout 300 python bench.py --no-extra --no-cpu-baseline --min-seconds 30 > $O/sustained.json 2> $O/sustained.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5u/sustained.json").read().strip().splitlines()[-1])
print(d["value"], d["sustained"])
PY
