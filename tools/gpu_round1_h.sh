#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
tail -6 gpurun_out/pytest_gpu_h.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants f0@0,f1@0,f1@32,f1@64,f1@96 > gpurun_out/bench_120mb_h.json 2> gpurun_out/bench_120mb_h.err
grep -E "variant|index" gpurun_out/bench_120mb_h.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0@0,f1@0,f1@16,f1@32,f1@64,f1@96,f0@64 > gpurun_out/bench_3100mb_h.json 2> gpurun_out/bench_3100mb_h.err
grep -E "variant|index" gpurun_out/bench_3100mb_h.err
cat gpurun_out/bench_3100mb_h.json
GSX_LOOKAHEAD=0 timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants f0@0,f0@32,f0@64,f0@96 > gpurun_out/bench_3100mb_h_nolook.json 2> gpurun_out/bench_3100mb_h_nolook.err
grep -E "variant|index" gpurun_out/bench_3100mb_h_nolook.err
