N=${1:-1}
echo "== full"; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --workload c3 --gpus $N --steps 5 --warmup 3 2> gpurun_out/c3_n$N.err | tee gpurun_out/c3_n$N.json | cut -c1-600; grep -i "error\|Traceback\|Maximum resident\|Elapsed" gpurun_out/c3_n$N.err | head; tail -3 gpurun_out/c3_n$N.err
