mkdir -p gpurun_out/r2a
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
timeout 900 python bench.py > gpurun_out/r2a/bench_default.json 2> gpurun_out/r2a/bench_default.err; echo "bench exit $?"
for mb in 4 5 6; do ZS_KLT_BLOCKS63=$mb python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > gpurun_out/r2a/bench_tumvi_mb$mb.json 2> gpurun_out/r2a/bench_tumvi_mb$mb.err; done
python bench.py --config TUMVI752 --no-extra --min-seconds 0 --no-cpu-baseline --steps 5 > gpurun_out/r2a/bench_tumvi752.json 2> gpurun_out/r2a/bench_tumvi752.err
ZS_KLT_NO_TMA=1 python bench.py --config TUMVI --no-extra --min-seconds 0 --no-cpu-baseline --steps 3 > gpurun_out/r2a/bench_tumvi_notma.json 2> gpurun_out/r2a/bench_tumvi_notma.err
tail -c 600 gpurun_out/r2a/bench_default.err
