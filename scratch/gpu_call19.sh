mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
timeout 300 python scratch/probe_kern.py "" "" 2>&1 | tail -2 | tee gpurun_out/r1c_probe_kern.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_step.csv python scratch/prof_step.py 4 > gpurun_out/r1c_ncu_launch.log 2>&1
python profiles/launch_summary.py gpurun_out/r1c_launches_step.csv | head -16
