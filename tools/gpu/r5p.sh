# round 2, session 4: ncu of the L2 tensor kernel with the tree drain, launch list of the L2 call
O=gpurun_out/r5p; mkdir -p $O
L="python tools/bench_l2.py --pairs 64 --reps 2"
timeout 300 $L > $O/plain.json 2> $O/plain.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_tc_persist -s 2 -c 1 -o $O/l2_tree $L > $O/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_l2.csv $L > $O/ncu_list.log 2>&1
ls -la $O
