#!/bin/bash
# second GPU session: builder parity, 120 Mb with GPU-built index, first 3.1 Gb (north-star size) numbers
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
tail -5 gpurun_out/pytest_gpu_b.log
timeout 900 python bench.py --genome-mb 120 --guides-per-step 20000 --steps 3 --warmup 3 --cpu-sample 4000 --sweep-variants 0,1,2,3,4 > gpurun_out/bench_120mb_b.json 2> gpurun_out/bench_120mb_b.err
tail -8 gpurun_out/bench_120mb_b.err
timeout 1500 python bench.py --genome-mb 3100 --n-chr 24 --seed 3 --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants 0,1,2,3,4 > gpurun_out/bench_3100mb_b.json 2> gpurun_out/bench_3100mb_b.err
tail -12 gpurun_out/bench_3100mb_b.err
cat gpurun_out/bench_3100mb_b.json
