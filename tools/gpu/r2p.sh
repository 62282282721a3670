# round 2, GPU call p: L2 per-call flag + sampled LK cache check: L2 / cache / adapter tests, seams bench
mkdir -p gpurun_out/r2p && O=gpurun_out/r2p
timeout 900 python -m pytest tests -m gpu -q -k "l2 or cache or adapter or mirror or errors" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 300 python tools/bench_seams.py > $O/seams.json 2>&1; tail -c 900 $O/seams.json
