#!/bin/bash
# Final-tree evidence of round 2: A/B of the small-grid collision CTAs, then full ncu captures of the C3 kernels
# (collision at 256 bins with 16-warp CTAs, cosine transforms, fused Thomas pass) on the slice scratch/prof_c3.py runs.
O=gpurun_out/prof3
mkdir -p $O
python scratch/probe_coll_small.py > $O/coll_small.log 2>&1; echo "probe rc=$?"; cat $O/coll_small.log
python scratch/prof_c3.py > $O/c3_plain.log 2>&1 || { cat $O/c3_plain.log; exit 1; }
cat $O/c3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_kernels.csv python scratch/prof_c3.py > $O/ncu_launch_c3.log 2>&1; echo "launch list c3 rc=$?"
for k in k_collide_struct k_dct_forward k_thomas_fused k_dct_inverse; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/full_c3_$k python scratch/prof_c3.py > $O/ncu_c3_$k.log 2>&1; echo "c3 $k rc=$?"
done
for f in $O/full_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
ncu -i $O/full_c3_k_collide_struct.ncu-rep --page source --csv > $O/full_c3_k_collide_struct_source.csv 2>/dev/null
rm -f $O/full_*.ncu-rep
ls -la $O
