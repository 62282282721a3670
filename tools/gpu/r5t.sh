# round 2, session 4: 4-GPU line of the last tree under torchrun
mkdir -p gpurun_out/r5t && O=gpurun_out/r5t
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r5t/bench_n4.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d.get("rank_ms_per_step"), (d["extra"].get("C4_sharded") or {}).get("value"))
PY
