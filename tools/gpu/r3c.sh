# round 2, GPU call: mirrors / adapter after the occupancy truncation fix (keypoints outside the frame)
mkdir -p gpurun_out/r3c && O=gpurun_out/r3c
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -8 $O/pytest.log
