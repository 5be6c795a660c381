mkdir -p gpurun_out
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N rc=$?"
cat gpurun_out/bench_n$N.json | cut -c1-400
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2>/dev/null; cut -c1-300 gpurun_out/bench_n1.json
