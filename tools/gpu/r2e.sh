# round 2, GPU call e: all GPU tests after the L2 / LK-cache / cell-check changes, L2 micro-bench + C3 line, launch list of the L2 call
mkdir -p gpurun_out/r2e && O=gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
timeout 300 python tools/bench_l2.py --pairs 64 > $O/l2_pairs64.json 2> $O/l2_pairs64.err; cat $O/l2_pairs64.json
timeout 300 python tools/bench_l2.py --pairs 1 > $O/l2_pairs1.json 2> $O/l2_pairs1.err
timeout 600 python bench.py --config C3 --steps 10 > $O/bench_C3.json 2> $O/bench_C3.err; tail -c 400 $O/bench_C3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_l2.csv python tools/bench_l2.py --pairs 64 --reps 2 > $O/ncu_l2.log 2>&1
