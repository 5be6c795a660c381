mkdir -p gpurun_out
timeout 600 python scratch/probe_kern.py "" "QPB_PIPE_REV=1" "QPB_PIPE_KEEPB=1" "QPB_PIPE_REV=1,QPB_PIPE_KEEPB=1" "" "QPB_PIPE_REV=1" 2>&1 | tail -6 | tee gpurun_out/r1c_probe_l2.log
