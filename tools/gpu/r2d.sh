# round 2, GPU call d: the C++ adapter harness + landmark tests
mkdir -p gpurun_out/r2d && O=gpurun_out/r2d
timeout 900 python -m pytest tests/test_gpu_adapter.py tests/test_gpu_landmarks.py -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -40 $O/pytest.log
