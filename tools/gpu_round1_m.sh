#!/bin/bash
# session M: sweep kernel with separate zero-budget pass; batch size and occupancy sweep
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "slice_major or sweep_kernel or fast_and_general" > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
tail -3 gpurun_out/pytest_gpu_m.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v0,s5v2,s5v3,s4v2,s6v2,s5v1 > gpurun_out/bench_3100mb_m.json 2> gpurun_out/bench_3100mb_m.err
grep -E "variant|index" gpurun_out/bench_3100mb_m.err
cat gpurun_out/bench_3100mb_m.json
timeout 1500 python bench.py --steps 2 --warmup 2 --guides-per-step 200000 --no-cpu-baseline --sweep-variants s5v0,s5v2,s4v2,s4v0 > gpurun_out/bench_3100mb_m200k.json 2> gpurun_out/bench_3100mb_m200k.err
grep -E "variant|index" gpurun_out/bench_3100mb_m200k.err
cat gpurun_out/bench_3100mb_m200k.json
