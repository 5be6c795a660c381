for L in r1 c88d7fc 25ce8c3 6cbb1b3; do echo "== $L"; QPB_LIB=scratch/libqpb_$L.so python scratch/probe_seg.py 2>&1 | grep -v "^library" | head -1; done
echo "== current"; python scratch/probe_seg.py 2>&1 | head -1
