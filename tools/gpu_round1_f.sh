#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f.log
tail -5 gpurun_out/pytest_gpu_f.log
timeout 1500 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
tail -4 gpurun_out/bench_f.err; cat gpurun_out/bench_f.json
