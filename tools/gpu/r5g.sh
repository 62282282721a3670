# round 2, session 4: ncu of the packed + rolled 63x63 KLT kernel, launch list of the TUMVI step
O=gpurun_out/r5g; mkdir -p $O
TV="python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 $TV > $O/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_klt_track_v4 -s 4 -c 1 -o $O/klt63_rolled $TV > $O/ncu_klt63.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_tumvi.csv $TV > $O/ncu_list.log 2>&1
ls -la $O
