#!/bin/bash
# Final round-1 evidence: plain run, launch list, full captures of the three hot kernels (one GPU).
mkdir -p gpurun_out
python scratch/prof_step.py 4 > gpurun_out/r1_plain.log 2>&1 || exit 1
cat gpurun_out/r1_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv python scratch/prof_step.py 4 > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_collide" -s 60 -c 3 -o gpurun_out/r1_hot python scratch/prof_step.py 4 > gpurun_out/r1_ncu_hot.log 2>&1
tail -2 gpurun_out/r1_ncu_hot.log
