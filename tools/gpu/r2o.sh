# round 2, GPU call o (8 GPUs): default bench under torchrun, 1 line with C4_sharded; then seams bench on rank 0's GPU
mkdir -p gpurun_out/r2o && O=gpurun_out/r2o
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "exit $?"; tail -c 400 $O/bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2o/bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["rank_ms_per_step"], d["sustained"]["value"], d["extra"]["C4_sharded"]["value"], d["extra"]["C4_sharded"].get("rank_ms_per_step"))
PY
timeout 300 python tools/bench_seams.py > $O/seams.json 2>&1; tail -c 600 $O/seams.json
