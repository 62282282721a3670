# round 2, late: ncu launch list of the default configuration on the final tree and a full capture of the packed-key Hamming kernel
O=gpurun_out/r4l; mkdir -p $O
C2="python bench.py --no-extra --no-cpu-baseline --min-seconds 0 --steps 2 --warmup 3"
timeout 300 $C2 > $O/plain_c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv $C2 > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hamming_top2 -s 4 -c 1 -o $O/hamming $C2 > $O/ncu_hamming.log 2>&1
ls -la $O
