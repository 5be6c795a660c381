#!/bin/bash
# Closing captures of the C3 kernels on the final tree (16-bin slice of the 2048^2 grid, scratch/prof_c3.py).
O=gpurun_out/prof2c
mkdir -p $O
python scratch/prof_c3.py > $O/c3_plain.log 2>&1 || exit 1
cat $O/c3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_kernels.csv python scratch/prof_c3.py > $O/ncu_launch_c3.log 2>&1; echo "launch list c3 rc=$?"
for k in k_collide_struct k_dct_forward k_thomas_frozen k_dct_inverse; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/full_c3_$k python scratch/prof_c3.py > $O/ncu_c3_$k.log 2>&1; echo "c3 $k rc=$?"
done
for f in $O/full_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
rm -f $O/full_*.ncu-rep
ls $O
