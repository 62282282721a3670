# round 2, GPU call r: compute-sanitizer (memcheck, then racecheck + synccheck on the kernels that share shared memory across
# lanes without CTA barriers) over small test subsets
mkdir -p gpurun_out/r2r && O=gpurun_out/r2r
export ZS_FE_NO_GRAPH=1
SEL="subpix or test_l2_u8 or test_l2_rejects or test_pyr_lk_mirror or test_keypoint_detector_parallel_mirror or test_match_temporal or test_device_tracker_empty_landmark or test_frontend_batches_vs_oracle"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > $O/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $O/memcheck.log; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/memcheck.log | tail -8
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "subpix or test_pyr_lk_mirror or test_l2_u8" > $O/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $O/racecheck.log; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/racecheck.log | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "subpix or test_pyr_lk_mirror" > $O/synccheck.log 2>&1; echo "synccheck exit $?" | tee -a $O/synccheck.log; grep -E "ERROR SUMMARY|passed|failed" $O/synccheck.log | tail -5
