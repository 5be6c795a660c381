mkdir -p gpurun_out
python scratch/prof_step.py 4 > gpurun_out/r1c_plain.log 2>&1 || exit 1
for k in k_sweep_x_pipe k_sweep_y_pipe; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/r1c_$k python scratch/prof_step.py 4 > gpurun_out/r1c_ncu_$k.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r1c_$k.ncu-rep --page source --csv > gpurun_out/r1c_${k}_source.csv 2>/dev/null
ncu -i gpurun_out/r1c_$k.ncu-rep --page raw --csv > gpurun_out/r1c_${k}_raw.csv 2>/dev/null
done
cat gpurun_out/r1c_plain.log
