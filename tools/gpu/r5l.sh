# round 2, session 4: 2-GPU line of the final tree under torchrun + the reference arm under torchrun
mkdir -p gpurun_out/r5l && O=gpurun_out/r5l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5l/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d.get("rank_ms_per_step"), (d["extra"].get("C4_sharded") or {}).get("value"))
PY
tail -2 $O/bench_n2.err
