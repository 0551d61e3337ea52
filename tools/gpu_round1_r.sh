#!/bin/bash
# session R: 16-row all-level summaries + work-unit parts; batch-size sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r.log
tail -3 gpurun_out/pytest_gpu_r.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v5,s5v3,s6v2,s6v5,s5v2p1,s5v2p2,s5v2p8,s4v2 > gpurun_out/bench_3100mb_r.json 2> gpurun_out/bench_3100mb_r.err
grep -E "variant|index" gpurun_out/bench_3100mb_r.err
cat gpurun_out/bench_3100mb_r.json
timeout 1500 python bench.py --steps 2 --warmup 2 --guides-per-step 200000 --no-cpu-baseline --sweep-variants s5v2,s5v5,s5v3,s6v5,s5v2p1,s5v2p2 > gpurun_out/bench_3100mb_r200k.json 2> gpurun_out/bench_3100mb_r200k.err
grep -E "variant|index" gpurun_out/bench_3100mb_r200k.err
cat gpurun_out/bench_3100mb_r200k.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_r python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r.log 2>&1
tail -3 gpurun_out/ncu_full_r.log
