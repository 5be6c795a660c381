#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
./scratch/fp64_micro > gpurun_out/fp64_micro.log 2>&1; cat gpurun_out/fp64_micro.log
timeout 600 python scratch/prof_e2e.py > gpurun_out/prof_e2e.log 2>&1; head -60 gpurun_out/prof_e2e.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
