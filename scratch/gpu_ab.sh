# A/B of the working tree's library against another build of it on one box: QPB_LIB_OLD=path bash scratch/gpu_ab.sh
for rep in 1 2; do
QPB_LIB=${QPB_LIB_OLD:-scratch/libqpb_prev.so} timeout 300 python scratch/probe_kern.py "" 2>&1 | tail -1
timeout 300 python scratch/probe_kern.py "" 2>&1 | tail -1
done
