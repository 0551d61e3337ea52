#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
tail -15 gpurun_out/pytest_gpu_g.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0,f1,f4 > gpurun_out/bench_120mb_g.json 2> gpurun_out/bench_120mb_g.err
grep -E "variant|index" gpurun_out/bench_120mb_g.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0,f1,f4 > gpurun_out/bench_3100mb_g.json 2> gpurun_out/bench_3100mb_g.err
grep -E "variant|index" gpurun_out/bench_3100mb_g.err
cat gpurun_out/bench_3100mb_g.json
GSX_LOOKAHEAD=0 timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3100mb_g_nolook.json 2> gpurun_out/bench_3100mb_g_nolook.err
cat gpurun_out/bench_3100mb_g_nolook.json
