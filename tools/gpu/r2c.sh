# round 2, GPU call c (re-run of a+b whose outputs were lost with the container): tests, benches, ncu launch lists + full captures
mkdir -p gpurun_out/r2c && O=gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi.json 2> $O/bench_tumvi.err
timeout 300 python bench.py --config TUMVI752 --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > $O/bench_tumvi752.json 2> $O/bench_tumvi752.err
ZS_KLT_NO_TMA=1 timeout 300 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 3 > $O/bench_tumvi_notma.json 2> $O/bench_tumvi_notma.err
C2="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
TV="python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 $C2 > $O/plain_c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv $C2 > $O/ncu_c2.log 2>&1
timeout 300 $TV > $O/plain_tv.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_tumvi.csv $TV > $O/ncu_tv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_klt_track_v4 -s 4 -c 1 -o $O/klt31 $C2 > $O/ncu_klt31.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_klt_track_v4 -s 4 -c 1 -o $O/klt63 $TV > $O/ncu_klt63.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_corner_subpix|k_fast_grid_v2" -s 8 -c 2 -o $O/tumvi_detect $TV > $O/ncu_det.log 2>&1
for r in klt31 klt63 tumvi_detect; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; done
ncu -i $O/klt31.ncu-rep --page source --csv > $O/klt31.source.csv 2>/dev/null
ls -la $O
