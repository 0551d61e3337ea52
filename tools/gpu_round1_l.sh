#!/bin/bash
# session L: ncu full capture of sweep_kernel at 3.1 Gb + rest of the parity suite
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
tail -5 gpurun_out/pytest_gpu_l.log
GSX_SWEEP_VARIANT=3 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_l python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_l.log 2>&1
tail -3 gpurun_out/ncu_full_l.log
