#!/bin/bash
# Round-1 evidence: plain bench run, ncu launch list of the same command, full captures of the hot kernels.
mkdir -p gpurun_out/prof
O=gpurun_out/prof
python bench.py --steps 20 --warmup 3 > $O/bench_plain.json 2> $O/bench_plain.err || exit 1
cat $O/bench_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scratch/prof_step.py 4 > $O/step_plain.log 2>&1 || exit 1
for k in k_collide_struct k_sweep_x_pipe k_sweep_y_pipe; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $O/full_$k python scratch/prof_step.py 4 > $O/ncu_$k.log 2>&1; echo "$k rc=$?"
done
python scratch/prof_gemm.py gemm > $O/gemm_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_collide_gemm -s 2 -c 1 -f -o $O/full_k_collide_gemm python scratch/prof_gemm.py gemm > $O/ncu_gemm.log 2>&1; echo "gemm rc=$?"
python scratch/prof_gemm.py seg > $O/seg_plain.log 2>&1 && for k in k_sweep_x_pipe k_sweep_y_pipe; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $O/full_seg_$k python scratch/prof_gemm.py seg > $O/ncu_seg_$k.log 2>&1; echo "seg $k rc=$?"
done
ls -la $O
