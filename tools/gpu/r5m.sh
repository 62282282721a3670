# round 2, session 4: L2 matcher drain with two min3 trees per 32-column chunk (ZS_L2_TREE) against the four-chain update
O=gpurun_out/r5m; mkdir -p $O
ZS_L2_TREE=1 timeout 600 python -m pytest tests -m gpu -x -q -k "l2" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python tools/bench_l2.py --pairs 64 > $O/l2_chains.json 2> $O/l2_chains.err
ZS_L2_TREE=1 timeout 300 python tools/bench_l2.py --pairs 64 > $O/l2_tree.json 2> $O/l2_tree.err
tail -n 3 $O/l2_chains.json; tail -n 3 $O/l2_tree.json
B="python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/c3_chains.json 2> $O/c3_chains.err
ZS_L2_TREE=1 timeout 300 $B > $O/c3_tree.json 2> $O/c3_tree.err
python - <<'PY'
import json
for f in ("c3_chains","c3_tree"):
    d=json.loads(open("gpurun_out/r5m/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["ms_per_step"], d["roofline"]["avg_launch_ms"], round(d["u8_rows"]["value"]))
PY
