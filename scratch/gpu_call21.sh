mkdir -p gpurun_out
timeout 300 python scratch/probe_kern.py "QPB_PIPE_PDL=0" "" "QPB_PIPE_PDL=0" "" 2>&1 | tail -4 | tee gpurun_out/r1c_probe_pdl.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
