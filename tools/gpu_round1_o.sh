#!/bin/bash
# session O: sweep kernel over the 32-row filter array -- parity, variants, ncu
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -3 gpurun_out/pytest_gpu_o.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v0,s5v2,s5v3,s4v2,s6v2,s5v4 > gpurun_out/bench_3100mb_o.json 2> gpurun_out/bench_3100mb_o.err
grep -E "variant|index" gpurun_out/bench_3100mb_o.err
cat gpurun_out/bench_3100mb_o.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_o python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_o.log 2>&1
tail -3 gpurun_out/ncu_full_o.log
