for cfg in "QPB_STAGED_D2H=0" "QPB_PREFAULT=0 QPB_COPY_THREADS=4" "QPB_PREFAULT=1 QPB_COPY_THREADS=4" "QPB_PREFAULT=1 QPB_COPY_THREADS=8" "QPB_PREFAULT=0 QPB_COPY_THREADS=8" "QPB_PREFAULT=1 QPB_COPY_THREADS=2"; do
echo "== $cfg"; env $cfg python scratch/probe_frames.py 2>&1 | tail -3
done | tee gpurun_out/r1c_frames.log
cat /sys/kernel/mm/transparent_hugepage/enabled; nproc
