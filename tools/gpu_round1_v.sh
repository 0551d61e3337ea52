#!/bin/bash
# session V: sweep kernel with per-guide uniform runs (masks in registers, unit-stride xor table)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v.log
tail -3 gpurun_out/pytest_gpu_v.log
timeout 1500 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --sweep-variants s4v2,s4v5,s4v3,s4v0,s5v2,s5v5,s3v2 > gpurun_out/bench_3100mb_v.json 2> gpurun_out/bench_3100mb_v.err
grep -E "variant|index" gpurun_out/bench_3100mb_v.err
cat gpurun_out/bench_3100mb_v.json
timeout 1500 python bench.py --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v5,s4v2,s5v3 > gpurun_out/bench_3100mb_v50k.json 2> gpurun_out/bench_3100mb_v50k.err
grep -E "variant|index" gpurun_out/bench_3100mb_v50k.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_v python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_v.log 2>&1
tail -3 gpurun_out/ncu_full_v.log
