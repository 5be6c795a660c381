mkdir -p gpurun_out
timeout 300 python scratch/probe_kern.py "" "" 2>&1 | tail -2 | tee gpurun_out/r1c_probe_bar.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
