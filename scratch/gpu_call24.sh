mkdir -p gpurun_out
timeout 300 python scratch/probe_kern.py "QPB_COLL_SPLIT=0" "" "QPB_COLL_CC=1" "QPB_COLL_SPLIT=0" "" 2>&1 | tail -5 | tee gpurun_out/r1c_probe_split.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
