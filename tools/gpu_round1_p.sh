#!/bin/bash
# session P: sweep kernel over pattern summaries -- parity, variants, ncu
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_p.log
tail -3 gpurun_out/pytest_gpu_p.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v0,s5v2,s5v3,s5v4,s4v2,s6v2,s6v3 > gpurun_out/bench_3100mb_p.json 2> gpurun_out/bench_3100mb_p.err
grep -E "variant|index" gpurun_out/bench_3100mb_p.err
cat gpurun_out/bench_3100mb_p.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_p python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_p.log 2>&1
tail -3 gpurun_out/ncu_full_p.log
