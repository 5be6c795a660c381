#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "sharded" > gpurun_out/pytest_sharded.log 2>&1; tail -3 gpurun_out/pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
cat gpurun_out/bench_n$N.json; tail -20 gpurun_out/bench_n$N.err
