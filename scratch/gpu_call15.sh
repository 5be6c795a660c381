mkdir -p gpurun_out
for rep in 1 2; do
QPB_LIB=scratch/libqpb_r1.so timeout 300 python scratch/probe_kern.py "" 2>&1 | tail -2
timeout 300 python scratch/probe_kern.py "" "QPB_COLL_TJ=8" 2>&1 | tail -2
done | tee gpurun_out/r1c_ab.log
