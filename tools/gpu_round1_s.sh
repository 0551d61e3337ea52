#!/bin/bash
# session S: sweep kernel with per-guide mask table in shared memory (hoisted plane masks, smem owner search)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s.log
tail -3 gpurun_out/pytest_gpu_s.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v5,s5v3,s5v0,s4v2,s4v5,s6v2,s5v2p2 > gpurun_out/bench_3100mb_s.json 2> gpurun_out/bench_3100mb_s.err
grep -E "variant|index" gpurun_out/bench_3100mb_s.err
cat gpurun_out/bench_3100mb_s.json
timeout 1500 python bench.py --steps 2 --warmup 2 --guides-per-step 200000 --no-cpu-baseline --sweep-variants s5v2,s5v5,s5v3,s4v2,s6v2 > gpurun_out/bench_3100mb_s200k.json 2> gpurun_out/bench_3100mb_s200k.err
grep -E "variant|index" gpurun_out/bench_3100mb_s200k.err
cat gpurun_out/bench_3100mb_s200k.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_s python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_s.log 2>&1
tail -3 gpurun_out/ncu_full_s.log
