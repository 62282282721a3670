# round 2, GPU call i: why the fused per-level pyramid builder loses to the three passes at 256 images (DESIGN "next (3)")
mkdir -p gpurun_out/r2i && O=gpurun_out/r2i
C2="python bench.py --no-extra --min-seconds 0 --no-cpu-baseline --steps 2 --warmup 3"
ZS_PYR_FUSED=1 timeout 300 $C2 > $O/bench_fused.json 2> $O/bench_fused.err
timeout 300 $C2 > $O/bench_split.json 2> $O/bench_split.err
python - <<'PY'
import json
for f in ("fused","split"):
    d=json.loads(open("gpurun_out/r2i/bench_%s.json"%f).read().strip().splitlines()[-1]); print(f, d["stage_ms_per_step"]["pyramid"])
PY
ZS_PYR_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pyr_level -s 8 -c 4 -o $O/pyr_fused $C2 > $O/ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pad_reflect|k_pyr_down|k_scharr" -s 20 -c 10 -o $O/pyr_split $C2 > $O/ncu_split.log 2>&1
ls -la $O
