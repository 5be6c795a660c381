timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 500 python scratch/probe_sizes.py coll 2>&1 | tee gpurun_out/probe_sizes_coll.log
