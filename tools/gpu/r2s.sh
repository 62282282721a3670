# round 2, GPU call s: 30-second sustained run of the default workload (clocks sampled throughout)
mkdir -p gpurun_out/r2s && O=gpurun_out/r2s
timeout 600 python bench.py --no-extra --no-cpu-baseline --min-seconds 30 > $O/bench_sustained30.json 2> $O/bench_sustained30.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s/bench_sustained30.json").read().strip().splitlines()[-1])
print(d["value"], d["sustained"])
PY
nvidia-smi --query-gpu=power.draw,clocks.sm,temperature.gpu --format=csv
