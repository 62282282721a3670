# round 2, session 4: the default line of the final tree (extras included), smoke, full GPU suite
O=gpurun_out/r5j; mkdir -p $O
( time timeout 900 python bench.py > $O/default.json 2> $O/default.err ) 2> $O/default.time; echo "bench exit $?"; tail -3 $O/default.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5j/default.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d.get("sustained",{}).get("value"))
x=d["extra"]
for k,v in x.items():
    if isinstance(v,dict): print(k, v.get("value"), (v.get("e2e") or {}).get("value") if isinstance(v.get("e2e"),dict) else None, v.get("error"))
PY
