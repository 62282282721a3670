# round 2, GPU call (4 GPUs): default bench under torchrun + reference arm under torchrun (rank 0 only works)
mkdir -p gpurun_out/r3b && O=gpurun_out/r3b
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $O/ref_n4.json 2> $O/ref_n4.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3b/bench_n4.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["rank_ms_per_step"], d["extra"]["C4_sharded"]["value"])
r=json.loads(open("gpurun_out/r3b/ref_n4.json").read().strip().splitlines()[-1]); print(r["impl"], r["value"], r["cpu_baseline"]["cores"])
PY
