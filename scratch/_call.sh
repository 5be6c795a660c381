QPB_LIB=scratch/libqpb_r1.so python scratch/probe_seg.py 2>&1 | grep -v "^library"
python scratch/probe_seg.py 2>&1
QPB_PIPE_PDL=0 python scratch/probe_seg.py 2>&1 | head -1
