#!/bin/bash
# C3 strong scaling, default bench command of the driver at N ranks (N = $1), results under gpurun_out/scale_final/
O=gpurun_out/scale_final
mkdir -p $O
n=$1
if [ "$n" = "1" ]; then
  python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "N=1 rc=$?"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "N=$n rc=$?"
fi
python - <<PY
import json
d=json.loads(open("$O/bench_n$n.json").read().strip().splitlines()[-1])
print($n, "ms/step", round(d["ms_per_step"],2), "value", d["value"], "e2e s", d["e2e"]["seconds"], "e2e value", d["e2e"]["value"], d["e2e"].get("breakdown_rank0_s"), {k:round(v,2) for k,v in d["time_shares"].items() if k!="note"})
PY
