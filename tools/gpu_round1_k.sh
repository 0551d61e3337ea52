#!/bin/bash
# session K: slice-major front end (sweep_kernel) -- parity + slice width / occupancy sweep at 3.1 Gb
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
tail -12 gpurun_out/pytest_gpu_k.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s0,s1v0,s2v0,s3v0,s3v2 > gpurun_out/bench_120mb_k.json 2> gpurun_out/bench_120mb_k.err
grep -E "variant|index" gpurun_out/bench_120mb_k.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s0,s4v0,s5v0,s6v0,s5v1,s5v2,s5v3,s5v4,s4v2,s6v2 > gpurun_out/bench_3100mb_k.json 2> gpurun_out/bench_3100mb_k.err
grep -E "variant|index" gpurun_out/bench_3100mb_k.err
cat gpurun_out/bench_3100mb_k.json
