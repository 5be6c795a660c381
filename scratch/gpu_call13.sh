set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r1c_pytest.log
timeout 300 python scratch/probe_kern.py "" "QPB_COLL_TJ=8" "" "QPB_COLL_TJ=8" 2>&1 | tail -6 | tee gpurun_out/r1c_probe_kern.log
