for x in 0 1; do
echo "== 2048 pipelined ne=16 extra halo $x"; QPB_HALO_EXTRA=$x QPB_DEBUG_RES=1 python scratch/probe_res.py 2048 16 2>&1 | grep "segments\|per bin" | tail -4
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 500 python scratch/probe_sizes.py diff 2>&1 | tee gpurun_out/probe_sizes_diff.log
