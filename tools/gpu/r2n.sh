# round 2, GPU call n: final build -- smoke, full GPU suite, default bench line (with extras), reference arm, launch list
mkdir -p gpurun_out/r2n && O=gpurun_out/r2n
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref exit $?"
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"; tail -c 300 $O/bench_default.err
C2="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 $C2 > $O/plain_c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv $C2 > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_klt_track_v4 -s 4 -c 1 -o $O/klt31 $C2 > $O/ncu_klt31.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["sustained"]["value"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline"].get("issue",{}).get("frac"))
print({k:(v.get("value") or v.get("stereo_frames_per_s")) if isinstance(v,dict) else v for k,v in d["extra"].items()})
PY
