mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
timeout 300 python scratch/prof_e2e.py 2>&1 | head -6 | tee gpurun_out/r1c_prof_e2e.log
