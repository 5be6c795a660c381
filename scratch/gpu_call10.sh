timeout 600 python -m pytest tests -m gpu -x -q -k "sharded" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config']['parallelism'], d['max_occupation'])"; tail -3 gpurun_out/bench_n2.err
QPB_NO_FUSED_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('a2a path', d['value'], d['ms_per_step'], d['max_occupation'])"
