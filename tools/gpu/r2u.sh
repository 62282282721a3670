# round 2, GPU call u: final build -- reference arm + default line (as the driver runs them)
mkdir -p gpurun_out/r2u && O=gpurun_out/r2u
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref exit $?"
timeout 1200 python bench.py --gpus 1 > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"; tail -c 300 $O/bench_default.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/r2u/bench_reference.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/r2u/bench_default.json").read().strip().splitlines()[-1])
print("ref", r["value"], "ours", d["value"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"], "ratio e2e/ref", d["e2e"]["value"]/r["value"])
print(d["stage_ms_per_step"]); print({k:(v.get("value") or v.get("stereo_frames_per_s")) if isinstance(v,dict) else v for k,v in d["extra"].items()})
print(d["roofline"]["frac"], d["roofline"]["issue"]["frac"], d["roofline"]["traffic"])
PY
