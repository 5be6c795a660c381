mkdir -p gpurun_out
for L in r1 c88d7fc 25ce8c3 6cbb1b3; do echo "== $L"; QPB_LIB=scratch/libqpb_$L.so python scratch/probe_seg.py 2>&1 | grep -v "^library" | head -1; done | tee gpurun_out/r1c_bisect.log
echo "== current"; python scratch/probe_seg.py 2>&1 | head -1 | tee -a gpurun_out/r1c_bisect.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1c_pytest.log
