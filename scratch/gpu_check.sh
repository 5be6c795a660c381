#!/bin/bash
# tests + bench (+ optional ncu launch list) on the GPU box; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
