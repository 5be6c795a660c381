mkdir -p gpurun_out
export QPB_PIPE_PREFETCH_DYN=1
for rep in 1 2; do
QPB_LIB=scratch/libqpb_r1.so timeout 300 python scratch/probe_kern.py "" 2>&1 | tail -1
timeout 300 python scratch/probe_kern.py "QPB_PIPE_PREFETCH=0" "QPB_PIPE_PREFETCH=1" 2>&1 | tail -2
done | tee gpurun_out/r1c_ab2.log
unset QPB_PIPE_PREFETCH_DYN
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
