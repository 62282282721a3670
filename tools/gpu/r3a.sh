# round 2, GPU call: adapter harness with processing + triangulation
mkdir -p gpurun_out/r3a && O=gpurun_out/r3a
timeout 900 python -m pytest tests/test_gpu_adapter.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -25 $O/pytest.log
