#!/bin/bash
O=gpurun_out/scale
mkdir -p $O
N=$1
for n in $N; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_c3_n$n.json 2> $O/bench_c3_n$n.err; echo "N=$n rc=$?"
  python -c "
import json; d=json.load(open('$O/bench_c3_n$n.json')); print($n, d['ms_per_step'], d['value'], {k:round(v,2) for k,v in d['time_shares'].items() if k!='note'}, d['e2e']['seconds'], d['parity']['collision']['n_max_rel'], d['parity']['diffusion_cn_equations']['max_norm_residual'])"
done
if [ "$2" = "c5" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scratch/run_c5.py > $O/c5_n8.json 2> $O/c5_n8.err; echo "c5 rc=$?"; cat $O/c5_n8.json
fi
