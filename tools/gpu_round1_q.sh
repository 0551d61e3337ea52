#!/bin/bash
# session Q: xor-table decode + prefetched work units; batch-size sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "slice_major or sweep_kernel or fast_and_general or golden" > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_q.log
tail -3 gpurun_out/pytest_gpu_q.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v3,s6v2,s6v3,s4v3 > gpurun_out/bench_3100mb_q.json 2> gpurun_out/bench_3100mb_q.err
grep -E "variant|index" gpurun_out/bench_3100mb_q.err
cat gpurun_out/bench_3100mb_q.json
timeout 1500 python bench.py --steps 2 --warmup 2 --guides-per-step 200000 --no-cpu-baseline --sweep-variants s5v2,s5v3,s6v2,s6v3,s4v3 > gpurun_out/bench_3100mb_q200k.json 2> gpurun_out/bench_3100mb_q200k.err
grep -E "variant|index" gpurun_out/bench_3100mb_q200k.err
cat gpurun_out/bench_3100mb_q200k.json
GSX_SWEEP_VARIANT=3 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_q python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_q.log 2>&1
tail -3 gpurun_out/ncu_full_q.log
