# round 2, GPU call z: final state -- smoke, full GPU suite, default line
mkdir -p gpurun_out/r2z && O=gpurun_out/r2z
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 1200 python bench.py --gpus 1 > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["sustained"]["value"], d["cpu_baseline"]["value"])
print({k:(v.get("value") or v.get("stereo_frames_per_s")) if isinstance(v,dict) else v for k,v in d["extra"].items()})
PY
