mkdir -p gpurun_out
python scratch/prof_step.py 4 > gpurun_out/r1c_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_collide_struct -s 2 -c 1 -f -o gpurun_out/r1c_coll python scratch/prof_step.py 4 > gpurun_out/r1c_ncu_coll.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r1c_coll.ncu-rep --page source --csv > gpurun_out/r1c_coll_source.csv 2>/dev/null
ncu -i gpurun_out/r1c_coll.ncu-rep --page raw --csv > gpurun_out/r1c_coll_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
