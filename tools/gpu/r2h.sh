# round 2, GPU call h (2 GPUs): default bench under torchrun with the C4_sharded extra
mkdir -p gpurun_out/r2h && O=gpurun_out/r2h
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?"; tail -c 600 $O/bench_n2.err; tail -c 1500 $O/bench_n2.json
