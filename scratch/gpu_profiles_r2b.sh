#!/bin/bash
# Round-2 closing evidence for the kernels added after gpu_profiles_r2.sh: plain runs first, then the ncu launch list of the
# same command and full captures (one GPU, each only after its command has exited 0 without ncu).
O=gpurun_out/prof2b
mkdir -p $O
python scratch/prof_step.py 4 > $O/step_plain.log 2>&1 || exit 1
cat $O/step_plain.log
python bench.py --workload c2 --steps 20 --warmup 3 > $O/bench_c2_plain.json 2> $O/bench_c2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 > $O/ncu_launch_c2.log 2>&1; echo "launch list c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_pr_resident -s 3 -c 1 -f -o $O/full_c2_k_pr_resident python scratch/prof_step.py 4 > $O/ncu_c2_k_pr_resident.log 2>&1; echo "c2 k_pr_resident rc=$?"
python scratch/probe_vard.py 256 128 > $O/vard_plain.log 2>&1 || exit 1
cat $O/vard_plain.log
for k in k_sweep_x_vard k_sweep_y_vard; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o $O/full_vard_$k python scratch/probe_vard.py 256 128 > $O/ncu_vard_$k.log 2>&1; echo "vard $k rc=$?"
done
for f in $O/full_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
ncu -i $O/full_c2_k_pr_resident.ncu-rep --page source --csv > $O/full_c2_k_pr_resident_source.csv 2>/dev/null
rm -f $O/full_*.ncu-rep
ls -la $O
