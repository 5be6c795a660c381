set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r1c_pytest.log
timeout 300 python scratch/probe_kern.py "" "QPB_COLL_TJ=8" "" "QPB_COLL_TJ=8" 2>&1 | tail -6 | tee gpurun_out/r1c_probe_kern.log
timeout 300 python scratch/prof_e2e.py 2>&1 | head -24 | tee gpurun_out/r1c_prof_e2e.log
QPB_STAGED_D2H=0 timeout 300 python scratch/prof_e2e.py 2>&1 | head -5 | tee gpurun_out/r1c_prof_e2e_plain.log
