#!/bin/bash
# Round-2 evidence: plain runs first, then the ncu launch list of the same commands and full captures of the hot kernels.
O=gpurun_out/prof2
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python bench.py --workload c2 --steps 20 --warmup 3 > $O/bench_c2_plain.json 2> $O/bench_c2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 > $O/ncu_launch_c2.log 2>&1; echo "launch list c2 rc=$?"
python scratch/prof_step.py 4 > $O/step_plain.log 2>&1 || exit 1
cat $O/step_plain.log
for k in k_collide_struct k_sweep_x_pipe k_sweep_y_pipe; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $O/full_c2_$k python scratch/prof_step.py 4 > $O/ncu_c2_$k.log 2>&1; echo "c2 $k rc=$?"
done
python scratch/prof_c3.py > $O/c3_plain.log 2>&1 || exit 1
cat $O/c3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_kernels.csv python scratch/prof_c3.py > $O/ncu_launch_c3.log 2>&1; echo "launch list c3 rc=$?"
for k in k_collide_struct k_dct_forward k_thomas_modes k_dct_inverse; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/full_c3_$k python scratch/prof_c3.py > $O/ncu_c3_$k.log 2>&1; echo "c3 $k rc=$?"
done
for f in $O/full_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
ncu -i $O/full_c3_k_collide_struct.ncu-rep --page source --csv > $O/full_c3_k_collide_struct_source.csv 2>/dev/null
ncu -i $O/full_c3_k_thomas_modes.ncu-rep --page source --csv > $O/full_c3_k_thomas_modes_source.csv 2>/dev/null
ncu -i $O/full_c3_k_dct_forward.ncu-rep --page source --csv > $O/full_c3_k_dct_forward_source.csv 2>/dev/null
rm -f $O/full_*.ncu-rep
ls -la $O
