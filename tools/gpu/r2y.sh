# round 2, GPU call y: random KLT sweeps over every launch / tile form (persistent forced on small problems)
mkdir -p gpurun_out/r2y && O=gpurun_out/r2y
timeout 1500 python -m pytest tests/test_gpu_random_sweep.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
