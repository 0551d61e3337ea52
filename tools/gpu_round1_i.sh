#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
tail -12 gpurun_out/pytest_gpu_i.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants f0,f1 > gpurun_out/bench_120mb_i.json 2> gpurun_out/bench_120mb_i.err
grep -E "variant|index" gpurun_out/bench_120mb_i.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants f0,f1 > gpurun_out/bench_3100mb_i.json 2> gpurun_out/bench_3100mb_i.err
grep -E "variant|index" gpurun_out/bench_3100mb_i.err
cat gpurun_out/bench_3100mb_i.json
for L in 13 12; do GSX_FTAB=$L timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3100mb_i_L$L.json 2> gpurun_out/bench_3100mb_i_L$L.err; cat gpurun_out/bench_3100mb_i_L$L.json | cut -c1-220; done
