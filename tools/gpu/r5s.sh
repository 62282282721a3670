# round 2, session 4: last tree -- default line, smoke, ncu launch list of C2 and a full capture of the Hamming tensor kernel
O=gpurun_out/r5s; mkdir -p $O
( time timeout 900 python bench.py > $O/default.json 2> $O/default.err ) 2> $O/default.time; echo "bench exit $?"; tail -3 $O/default.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
C2="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv $C2 > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_tc_persist -s 4 -c 1 -o $O/hamming_tc $C2 > $O/ncu_hamming.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/reference.json 2> $O/reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5s/default.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d.get("sustained",{}).get("value"), d["stage_ms_per_step"])
for k,v in d["extra"].items():
    if isinstance(v,dict): print(k, v.get("value"), v.get("error"))
r=json.loads(open("gpurun_out/r5s/reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
ls $O
