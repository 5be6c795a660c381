#!/bin/bash
# Round-1 closing evidence: GPU tests, the bench line, the ncu launch list of the same command, full captures of the hot kernels.
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err || { tail -5 $O/bench.err; exit 1; }
cat $O/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scratch/prof_step.py 4 > $O/step_plain.log 2>&1 || exit 1
for k in k_collide_struct k_sweep_x_pipe k_sweep_y_pipe; do
  skip=40; [ $k = k_collide_struct ] && skip=2   # 8 collision launches in the 4 profiled steps, 144 sweeps
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $O/full_$k python scratch/prof_step.py 4 > $O/ncu_$k.log 2>&1; echo "$k rc=$?"
  ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/full_${k}_raw.csv 2>/dev/null
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la $O | head -30
